"""World-size-2 tests of the multi-GPU host logic on CPU ranks (gloo): the
partition of a map by spatial cell, the arg-min exchange of packed keys and the
sharding of candidate pairs (mola-fe-lidar_b200/multi_gpu.py).  The
per-shard search / merge kernels are replaced by an oracle-based stand-in here;
the CUDA path is covered by tests/test_gpu_multi.py."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleShardSearch:
    """CPU stand-in of CudaShardSearch: oracle kNN per shard, numpy merge."""

    def __init__(self, O, shard_xyz, global_index):
        import torch
        self.O, self.torch = O, torch
        self.cloud = O.Cloud(shard_xyz) if len(shard_xyz) else None
        self.gidx = np.asarray(global_index, dtype=np.uint32)

    def partial_keys(self, queries, k, max_dist):
        from mola_fe_lidar_b200 import multi_gpu as M
        nq = len(queries)
        if self.cloud is None:
            keys = np.full((nq, k), M.NO_KEY, dtype=np.uint64)
        else:
            cap = np.float32(max_dist) * np.float32(max_dist)
            idx, d2 = self.O.knn(self.cloud, queries, k, cap, kdtree=True)
            g = np.where(idx == 0xFFFFFFFF, 0xFFFFFFFF, self.gidx[np.minimum(idx, len(self.gidx) - 1)])
            keys = M.pack_keys(d2, g)
        return self.torch.from_numpy(keys.view(np.int64).copy())

    def merge(self, parts):
        from mola_fe_lidar_b200 import multi_gpu as M
        P, n, k = parts.shape
        m = M.merge_keys_numpy(parts.numpy().view(np.uint64), k)
        return self.torch.from_numpy(m.view(np.int64).copy())


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    import torch.distributed as dist
    import oracle_api as O
    from mola_fe_lidar_b200 import multi_gpu as M
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(7)  # same data on every rank
        themap = rng.uniform([-20, -20, -2], [20, 20, 2], size=(6000, 3)).astype(np.float32)
        themap[100:110] = themap[90:100]  # exact duplicates: ties broken by the lower global index
        q = rng.uniform([-21, -21, -2], [21, 21, 2], size=(500, 3)).astype(np.float32)
        q[:5] = themap[100:105]
        owner = M.partition_by_cell(themap, world, cell=4.0)
        mine = M.shard_indices(owner, rank)
        sm = M.ShardedMap(OracleShardSearch(O, themap[mine], mine), rank, world, dist)
        res = {}
        for k, r in ((1, 1.5), (6, 1.5), (6, 0.3)):
            keys = sm.query(q, k, r).numpy().view(np.uint64)
            res[f"k{k}_r{r}"] = keys
        # candidate pairs: i -> rank i mod world, gathered in pair order
        n_pairs = 7
        mine_pairs = M.pairs_of_rank(n_pairs, rank, world)
        local = np.array([[p, 10.0 * p + 1, rank] for p in mine_pairs], dtype=np.float64).reshape(-1, 3)
        res["pairs"] = M.gather_pair_results(local, n_pairs, rank, world, dist)
        res["owner"] = owner
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **res)
    finally:
        dist.destroy_process_group()


def test_partition_is_balanced_compact_and_order_preserving():
    sys.path.insert(0, ROOT)
    from mola_fe_lidar_b200 import multi_gpu as M
    rng = np.random.default_rng(3)
    pts = rng.uniform([-50, -50, -2], [50, 50, 2], size=(20000, 3)).astype(np.float32)
    for world in (1, 2, 4, 8):
        owner = M.partition_by_cell(pts, world, cell=8.0)
        assert owner.min() == 0 and owner.max() == world - 1
        counts = np.bincount(owner, minlength=world)
        assert counts.max() < 1.5 * len(pts) / world + 500  # cells are atomic, so only roughly equal
        parts = [M.shard_indices(owner, r) for r in range(world)]
        assert sum(len(p) for p in parts) == len(pts)
        for p in parts:
            assert np.all(np.diff(p.astype(np.int64)) > 0)  # increasing: ties order alike
        # a coarse cell never straddles two ranks
        cells = np.floor((pts[:, :2] - pts[:, :2].min(axis=0)) / 8.0).astype(np.int64)
        cid = cells[:, 0] * 100000 + cells[:, 1]
        for c in np.unique(cid)[:50]:
            assert len(np.unique(owner[cid == c])) == 1
    assert len(M.partition_by_cell(np.zeros((0, 3), np.float32), 4)) == 0


def test_pack_unpack_and_merge_keys():
    sys.path.insert(0, ROOT)
    from mola_fe_lidar_b200 import multi_gpu as M
    d2 = np.array([[0.0, 0.25, np.inf], [1.5, 1.5, 2.0]], dtype=np.float32)
    idx = np.array([[5, 9, 0xFFFFFFFF], [7, 3, 1]], dtype=np.uint32)
    keys = M.pack_keys(d2, idx)
    assert keys[0, 2] == M.NO_KEY
    i2, e2 = M.unpack_keys(keys)
    assert np.array_equal(i2, idx) and np.array_equal(e2, d2)
    # integer order of the key = (d2, index) order
    assert keys[1, 1] < keys[1, 0] < keys[1, 2]
    a = np.sort(M.pack_keys(np.float32([[0.1, 0.4, 0.9]]), np.uint32([[1, 2, 3]])), axis=1)
    b = np.sort(M.pack_keys(np.float32([[0.1, 0.2, np.inf]]), np.uint32([[0, 8, 0xFFFFFFFF]])), axis=1)
    m = M.merge_keys_numpy(np.stack([a, b]), 3)
    mi, md = M.unpack_keys(m)
    assert mi.tolist() == [[0, 1, 8]] and np.allclose(md, [[0.1, 0.1, 0.2]])


@pytest.mark.timeout(300)
def test_sharded_map_and_pair_sharding_world2(tmp_path, oracle):
    import torch.multiprocessing as mp
    import socket
    with socket.socket() as sk:  # a free port: suites may run side by side
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    world = 2
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0 = np.load(tmp_path / "rank0.npz")
    r1 = np.load(tmp_path / "rank1.npz")
    from mola_fe_lidar_b200 import multi_gpu as M
    rng = np.random.default_rng(7)
    themap = rng.uniform([-20, -20, -2], [20, 20, 2], size=(6000, 3)).astype(np.float32)
    themap[100:110] = themap[90:100]
    q = rng.uniform([-21, -21, -2], [21, 21, 2], size=(500, 3)).astype(np.float32)
    q[:5] = themap[100:105]
    full = oracle.Cloud(themap)
    for k, r in ((1, 1.5), (6, 1.5), (6, 0.3)):
        name = f"k{k}_r{r}"
        assert np.array_equal(r0[name], r1[name]), "ranks disagree"
        idx, d2 = oracle.knn(full, q, k, np.float32(r) * np.float32(r), kdtree=True)
        gi, gd = M.unpack_keys(r0[name])
        assert np.array_equal(gi, idx), name  # bit-exact vs the unsharded search, ties included
        assert np.array_equal(gd, d2), name
    # the duplicated points: the lower global index (90..94) wins the tie at d2 = 0
    gi, gd = M.unpack_keys(r0["k1_r1.5"])
    assert gi[:5, 0].tolist() == [90, 91, 92, 93, 94] and np.all(gd[:5, 0] == 0)
    assert len(np.unique(r0["owner"])) == 2
    for r in (r0, r1):
        assert np.array_equal(r["pairs"][:, 0], np.arange(7))
        assert np.array_equal(r["pairs"][:, 1], 10.0 * np.arange(7) + 1)
        assert np.array_equal(r["pairs"][:, 2], np.arange(7) % 2)
