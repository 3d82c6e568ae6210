"""multi_gpu.py -- the two places where the ICP path shards over the GPUs of one
box, one process per GPU over torch.distributed:

1. Batches of independent registrations (the reference's worker_pool_past_KFs_
   jobs, LidarOdometry.cpp:711-729, and the Monte-Carlo loop, cpp:775-787):
   candidate pair i -> rank i mod world, all Monte-Carlo samples of a pair on
   the same rank (they share the pair's cloud indices); no collective during
   compute, one all-gather of the result records at the end.

2. A local map too large or too hot for one GPU, split by spatial cell:
   every rank indexes its own cells, all ranks search all queries against
   their shard, and the per-rank partial arg-min lists are merged over NVLink.
   A partial result is a packed key (d2 as float32 bits) << 32 | global index:
   its integer order IS the (d2, index) order of the tie rule (Appendix A.4),
   so k = 1 merges with ONE all-reduce(MIN) on int64 and k > 1 with one
   all-gather of the per-query lists followed by the k-way merge kernel
   (b200icp_merge_keys_device).  Results equal the unsharded search bit for
   bit.

The search / merge kernels are injected (`ShardSearch`): CudaShardSearch is the
product path (C ABI, device tensors, NCCL); tests on CPU ranks (gloo) inject a
stand-in built on the oracle to exercise the partition and the exchange.
"""
import numpy as np

NO_KEY = 0x7F800000FFFFFFFF  # B200ICP_NO_KEY: +inf distance, invalid index


# ------------------------------------------------------------------ batches
def pairs_of_rank(n_pairs, rank, world):
    """pair i -> rank i mod world."""
    return list(range(rank, n_pairs, world))


def gather_pair_results(local, n_pairs, rank, world, dist, device=None):
    """local: float64 array [len(pairs_of_rank), F] of this rank's result
    records -> [n_pairs, F] in pair order on every rank (one all-gather)."""
    import torch
    local = np.ascontiguousarray(local, dtype=np.float64)
    per = (n_pairs + world - 1) // world
    F = local.shape[1] if local.ndim == 2 else 0
    buf = torch.zeros((per, F), dtype=torch.float64, device=device)
    if len(local):
        buf[:len(local)] = torch.from_numpy(local).to(buf.device)
    if world == 1:
        return buf[:n_pairs].cpu().numpy()
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    out = np.zeros((n_pairs, F))
    for r in range(world):
        idx = pairs_of_rank(n_pairs, r, world)
        out[idx] = parts[r][:len(idx)].cpu().numpy()
    return out


# --------------------------------------------------------------- sharded map
def _spread16(v):
    v = v.astype(np.uint64) & np.uint64(0xFFFF)
    v = (v | (v << np.uint64(8))) & np.uint64(0x00FF00FF)
    v = (v | (v << np.uint64(4))) & np.uint64(0x0F0F0F0F)
    v = (v | (v << np.uint64(2))) & np.uint64(0x33333333)
    v = (v | (v << np.uint64(1))) & np.uint64(0x55555555)
    return v


def partition_by_cell(points, world, cell=8.0, mode="blocks"):
    """Owner rank of every map point.  Coarse (x, y) cells of edge `cell` are
    ordered along a Morton curve and either cut into `world` contiguous runs of
    about equal point count ("blocks": every rank owns a compact set of cells)
    or dealt round-robin ("interleaved": every rank holds 1/world of every
    neighbourhood, which balances the load of a scan that only overlaps a
    small part of the map).  Deterministic from the points alone, so every
    rank computes the same partition."""
    pts = np.asarray(points, dtype=np.float32).reshape(-1, 3)
    if len(pts) == 0:
        return np.zeros(0, dtype=np.int32)
    finite = np.isfinite(pts).all(axis=1)
    lo = pts[finite, :2].min(axis=0) if finite.any() else np.zeros(2, np.float32)
    c = np.zeros((len(pts), 2), dtype=np.int64)
    c[finite] = np.floor((pts[finite, :2] - lo) / np.float32(cell)).astype(np.int64)
    c = np.clip(c, 0, 0xFFFF)
    code = _spread16(c[:, 0]) | (_spread16(c[:, 1]) << np.uint64(1))
    uniq, inv, cnt = np.unique(code, return_inverse=True, return_counts=True)
    if mode == "interleaved":
        owner_of_cell = (np.arange(len(uniq)) % world).astype(np.int32)
    elif mode == "blocks":
        before = np.cumsum(cnt) - cnt  # points in the cells ahead on the curve
        owner_of_cell = np.minimum(world - 1, (before * world) // max(len(pts), 1)).astype(np.int32)
    else:
        raise ValueError("mode must be 'blocks' or 'interleaved'")
    return owner_of_cell[inv]


def shard_indices(owner, rank):
    """Global indices of the rank's points, ascending: shard-local order keeps
    the global order, so ties break alike in both numberings."""
    return np.nonzero(np.asarray(owner) == rank)[0].astype(np.uint32)


def pack_keys(d2, idx):
    d2 = np.ascontiguousarray(d2, dtype=np.float32)
    k = (d2.view(np.uint32).astype(np.uint64) << np.uint64(32)) | np.asarray(idx, dtype=np.uint64)
    return np.where(np.asarray(idx) == 0xFFFFFFFF, np.uint64(NO_KEY), k)


def unpack_keys(keys):
    """packed keys -> (idx uint32, d2 float32); NO_KEY -> (0xFFFFFFFF, inf)."""
    keys = np.asarray(keys).astype(np.uint64)
    idx = (keys & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    d2 = (keys >> np.uint64(32)).astype(np.uint32).view(np.float32)
    return idx, d2


def merge_keys_numpy(parts, k):
    """[P, nq, k] ascending lists -> [nq, k]: the k smallest of the union
    (definition of b200icp_merge_keys_device; used by the CPU test stand-in)."""
    parts = np.asarray(parts).astype(np.uint64)
    P, nq, kk = parts.shape
    allk = np.sort(np.transpose(parts, (1, 0, 2)).reshape(nq, P * kk), axis=1)
    return allk[:, :k]


class CudaShardSearch:
    """Product path: this rank's shard as an indexed cloud on its GPU, partial
    keys and merges by the library's kernels, tensors stay on the device."""

    def __init__(self, icp, shard_xyz, global_index, search_radius, device):
        import torch
        self.icp, self.torch, self.device = icp, torch, device
        self.cloud = icp.upload(np.asarray(shard_xyz, dtype=np.float32), search_radius=search_radius)
        self.index_map = torch.from_numpy(np.asarray(global_index, dtype=np.uint32).view(np.int32).copy()).to(device)

    def upload_queries(self, xyz, search_radius):
        return self.icp.upload(np.asarray(xyz, dtype=np.float32), search_radius=search_radius)

    def partial_keys(self, queries, k, max_dist):
        nq = len(queries)
        out = self.torch.empty((nq, k), dtype=self.torch.int64, device=self.device)
        if nq:
            self.icp.knn_keys_device(self.cloud, queries, k, max_dist, out.data_ptr(),
                                     self.index_map.data_ptr() if self.index_map.numel() else 0)
        return out

    def merge(self, parts):
        P, n, k = parts.shape
        out = self.torch.empty((n, k), dtype=self.torch.int64, device=self.device)
        if n:
            self.torch.cuda.current_stream().synchronize()  # the exchange has landed
            self.icp.merge_keys_device(parts.data_ptr(), P, n * k, n, k, out.data_ptr())
        return out

    def close(self):
        self.cloud.free()


class PeerExchange:
    """Exchange buffers of the fused search + arg-min exchange: every rank owns
    one device buffer that all other ranks of the node map through CUDA IPC, so
    that the search kernel itself stores its rows into every rank's buffer over
    NVLink (b200icp_knn_keys_scatter)."""

    HEADER = 256  # bytes in front of the data: the barrier flags (one uint64 per rank)

    def __init__(self, icp, rank, world, dist, nbytes):
        self.icp, self.rank, self.world, self.nbytes = icp, rank, world, nbytes
        self.base, handle = icp.peer_alloc(nbytes + self.HEADER)  # zero-initialised: flags start at 0
        handles = [None] * world
        if world > 1:
            dist.all_gather_object(handles, handle)  # also orders the zeroing before any remote access
        else:
            handles[0] = handle
        self.bases = []
        try:
            for r in range(world):
                self.bases.append(self.base if r == rank else icp.peer_open(handles[r]))
        except Exception:
            self.close()  # unmap what was mapped, free the local buffer
            raise
        self.ptrs = [b + self.HEADER for b in self.bases]
        self.local = self.ptrs[rank]
        self.epoch = 0

    def barrier(self):
        """all ranks of the node, through the flags in peer memory"""
        self.epoch += 1
        self.icp.peer_barrier(self.bases, self.rank, self.epoch)

    def close(self):
        for r, p in enumerate(getattr(self, "bases", [])):
            if r != self.rank and p:
                self.icp.peer_close(p)
        self.bases, self.ptrs = [], []
        if getattr(self, "base", None):
            self.icp.peer_free(self.base)
            self.base = self.local = None


class ShardedMap:
    """One rank's handle on a spatially sharded map: query() returns the same
    [nq, k] packed keys on every rank.  query() exchanges with NCCL collectives;
    query_fused() lets the search kernel write into all ranks' buffers itself
    (peer memory over NVLink) and needs only barriers."""

    def __init__(self, search, rank, world, dist=None):
        self.search, self.rank, self.world, self.dist = search, rank, world, dist
        self.last_exchange_bytes = 0
        self.peers = None

    def _barrier(self):
        if self.world > 1:
            self.dist.barrier()

    def query_fused(self, queries, k, max_dist):
        """CudaShardSearch only.  The kernels store straight into the ranks'
        exchange buffers over NVLink.  k = 1: every rank folds its key into the
        slot of the query on the rank that owns the query (system-scope
        atomicMin: search + reduce-scatter in one kernel), the owner stores its
        slice into every rank's buffer; k > 1: the row of a query goes to its
        owner, the owner merges its slice and stores the merged rows into every
        rank's buffer (merge + all-gather in one kernel)."""
        import torch
        s, nq = self.search, len(queries)
        per = (nq + self.world - 1) // self.world
        need = max(8, 2 * self.world * per * k * 8)  # gather region + result region (see b200icp_knn_keys_exchange)
        if self.peers is None or self.peers.nbytes < need:
            # no collective here: nobody touches a rank's buffer after the last barrier of a query, and a
            # rank whose exchange could not be set up must not leave the others waiting in one
            if self.peers is not None:
                self.peers.close()
                self.peers = None
            self.peers = PeerExchange(s.icp, self.rank, self.world, self.dist, need)
        px = self.peers
        out = torch.empty((nq, k), dtype=torch.int64, device=s.device)
        if nq == 0:
            return out
        # reset -> barrier -> search + scatter -> barrier -> merge: one library call, one stream
        px.epoch = s.icp.knn_keys_exchange(s.cloud, queries, k, max_dist, px.bases, self.rank, px.epoch,
                                           out.data_ptr(), s.index_map.data_ptr() if s.index_map.numel() else 0)
        self.last_exchange_bytes = nq * k * 8 * (self.world - 1)
        return out

    def close(self):
        if self.peers is not None:  # safe without a collective: every query ends with a barrier
            self.peers.close()
            self.peers = None

    def query(self, queries, k, max_dist):
        import torch
        keys = self.search.partial_keys(queries, k, max_dist)  # [nq, k] int64
        self.last_exchange_bytes = 0
        if self.world == 1:
            return keys
        dist = self.dist
        if k == 1:
            # keys are non-negative as int64, so MIN is the unsigned (d2, index) minimum
            dist.all_reduce(keys, op=dist.ReduceOp.MIN)
            self.last_exchange_bytes = keys.numel() * 8
            return keys
        # every rank gathers every rank's per-query lists (one all-gather over
        # NVLink) and merges them with the k-way merge kernel
        nq = keys.shape[0]
        parts = torch.empty((self.world, nq, k), dtype=torch.int64, device=keys.device)
        if hasattr(dist, "all_gather_into_tensor") and keys.is_cuda:
            dist.all_gather_into_tensor(parts, keys)
        else:  # gloo (CPU test ranks)
            lst = [torch.empty_like(keys) for _ in range(self.world)]
            dist.all_gather(lst, keys)
            parts = torch.stack(lst)
        self.last_exchange_bytes = parts.numel() * 8
        return self.search.merge(parts)
