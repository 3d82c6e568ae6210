"""GPU parity of the sharded-map entry points (b200icp_knn_keys_device,
b200icp_merge_keys_device) through the C ABI: a map split by spatial cell into
P shards on ONE GPU, partial keys per shard, merged -- against the oracle's
unsharded search, bit-exact (world-size-2 plumbing is covered on CPU
ranks in test_multi_gpu_cpu.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _scene(rng, n_map=40000, n_q=5000):
    themap = rng.uniform([-30, -30, -2], [30, 30, 2], size=(n_map, 3)).astype(np.float32)
    themap[200:220] = themap[100:120]  # exact duplicates: ties
    q = rng.uniform([-31, -31, -2], [31, 31, 2], size=(n_q, 3)).astype(np.float32)
    q[:10] = themap[200:210]
    return themap, q


@pytest.mark.parametrize("k,radius", [(1, 1.0), (6, 1.0), (6, 0.35), (8, 1.0)])
@pytest.mark.parametrize("parts,mode", [(1, "blocks"), (3, "blocks"), (4, "interleaved")])
def test_sharded_knn_matches_unsharded_oracle(icp, oracle, k, radius, parts, mode):
    import torch
    from mola_fe_lidar_b200 import multi_gpu as M
    rng = np.random.default_rng(11)
    themap, q = _scene(rng)
    dev = torch.device("cuda", 0)
    owner = M.partition_by_cell(themap, parts, cell=5.0, mode=mode)
    qc = icp.upload(q, search_radius=radius)
    allkeys = torch.empty((parts, len(q), k), dtype=torch.int64, device=dev)
    shards = []
    for p in range(parts):
        mine = M.shard_indices(owner, p)
        s = M.CudaShardSearch(icp, themap[mine], mine, radius, dev)
        shards.append(s)
        allkeys[p] = s.partial_keys(qc, k, radius)
    merged = shards[0].merge(allkeys)
    torch.cuda.synchronize()
    gi, gd = M.unpack_keys(merged.cpu().numpy().view(np.uint64))
    idx, d2 = oracle.knn(oracle.Cloud(themap), q, k, np.float32(radius) * np.float32(radius), kdtree=True)
    assert np.array_equal(gi, idx)
    assert np.array_equal(gd, d2)
    if k == 1:  # what the all-reduce(MIN) over ranks computes
        mn = allkeys.min(dim=0).values
        assert torch.equal(mn, merged)
    for s in shards:
        s.close()
    qc.free()


def test_keys_device_equals_host_knn_and_handles_empty(icp):
    import torch
    from mola_fe_lidar_b200 import multi_gpu as M
    rng = np.random.default_rng(5)
    themap, q = _scene(rng, 20000, 3000)
    q[7] = np.nan  # a non-finite query has no neighbours
    dev = torch.device("cuda", 0)
    ref, qc = icp.upload(themap, search_radius=0.8), icp.upload(q, search_radius=0.8)
    keys = torch.empty((len(q), 6), dtype=torch.int64, device=dev)
    icp.knn_keys_device(ref, qc, 6, 0.8, keys.data_ptr())
    gi, gd = M.unpack_keys(keys.cpu().numpy().view(np.uint64))
    idx, d2 = icp.knn(ref, qc, 6, 0.8)
    assert np.array_equal(gi, idx) and np.array_equal(gd, d2)
    assert np.all(gi[7] == 0xFFFFFFFF) and np.all(np.isinf(gd[7]))
    # an empty shard contributes nothing
    empty = M.CudaShardSearch(icp, np.zeros((0, 3), np.float32), np.zeros(0, np.uint32), 0.8, dev)
    ek = empty.partial_keys(qc, 6, 0.8)
    assert bool((ek == M.NO_KEY).all())
    empty.close(), ref.free(), qc.free()


def test_pair_sharding_single_rank_gather():
    from mola_fe_lidar_b200 import multi_gpu as M
    rec = np.arange(12, dtype=np.float64).reshape(4, 3)
    out = M.gather_pair_results(rec, 4, 0, 1, None)
    assert np.array_equal(out, rec)


@pytest.mark.parametrize("k", [1, 6])
def test_fused_scatter_matches_unsharded_oracle(icp, oracle, k):
    """b200icp_knn_keys_scatter with two virtual ranks on one GPU: each rank's
    search writes into BOTH gather buffers (plain device buffers stand in for
    the IPC-mapped peers); k = 1 folds with atomicMin, k = 6 merges afterwards."""
    import torch
    from mola_fe_lidar_b200 import multi_gpu as M
    rng = np.random.default_rng(21)
    themap, q = _scene(rng, 30000, 4000)
    q[3] = np.inf
    radius, world = 1.0, 2
    dev = torch.device("cuda", 0)
    owner = M.partition_by_cell(themap, world, cell=5.0, mode="interleaved")
    qc = icp.upload(q, search_radius=radius)
    nq = len(q)
    bufs = [torch.full((world * nq * k,), M.NO_KEY, dtype=torch.int64, device=dev) for _ in range(world)]
    if k > 1:
        for b in bufs:
            b.fill_(12345)  # stale data: the call must overwrite every row of its slice
    torch.cuda.synchronize()
    shards = []
    for r in range(world):
        mine = M.shard_indices(owner, r)
        s = M.CudaShardSearch(icp, themap[mine], mine, radius, dev)
        shards.append(s)
        icp.knn_keys_scatter(s.cloud, qc, k, radius, [b.data_ptr() for b in bufs], r, s.index_map.data_ptr(),
                             atomic_min=(k == 1))
    torch.cuda.synchronize()
    idx, d2 = oracle.knn(oracle.Cloud(themap), q, k, np.float32(radius) * np.float32(radius), kdtree=True)
    for b in bufs:  # every "rank" ends up with the same data
        if k == 1:
            merged = b[:nq].reshape(nq, 1)
        else:
            merged = shards[0].merge(b.reshape(world, nq, k))
        gi, gd = M.unpack_keys(merged.cpu().numpy().view(np.uint64))
        assert np.array_equal(gi, idx) and np.array_equal(gd, d2)
        assert np.all(gi[3] == 0xFFFFFFFF)
    for s in shards:
        s.close()
    qc.free()


@pytest.mark.parametrize("k", [1, 6])
def test_one_call_exchange_single_rank(icp, oracle, k):
    """b200icp_knn_keys_exchange with world = 1 (exchange buffer from
    b200icp_peer_alloc, barrier flags in its header): reset, barrier, search +
    scatter, barrier and merge in one call; twice, to exercise the epochs."""
    import torch
    from mola_fe_lidar_b200 import multi_gpu as M
    rng = np.random.default_rng(33)
    themap, q = _scene(rng, 20000, 3000)
    dev = torch.device("cuda", 0)
    idx_map = np.arange(len(themap), dtype=np.uint32)
    search = M.CudaShardSearch(icp, themap, idx_map, 0.9, dev)
    sm = M.ShardedMap(search, 0, 1, None)
    qc = icp.upload(q, search_radius=0.9)
    idx, d2 = oracle.knn(oracle.Cloud(themap), q, k, np.float32(0.9) * np.float32(0.9), kdtree=True)
    for _ in range(2):
        keys = sm.query_fused(qc, k, 0.9)
        gi, gd = M.unpack_keys(keys.cpu().numpy().view(np.uint64))
        assert np.array_equal(gi, idx) and np.array_equal(gd, d2)
    assert sm.peers.epoch == 6  # three barriers per query
    assert torch.equal(keys, sm.query(qc, k, 0.9))
    sm.close(), qc.free(), search.close()
