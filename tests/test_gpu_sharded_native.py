"""The library's own multi-GPU host (b200icp_comm_*, b200icp_sharded_*; csrc/sharded.inl): a map split by
spatial cell, the NCCL communicator owned by the C++ side, a registration against the sharded map that is
BIT-IDENTICAL to b200icp_align against the unsharded one (BASELINE config 5).
World size 1 runs in-process on any GPU box; world size 2 needs two GPUs (NCCL refuses two ranks on one
device) and is skipped otherwise -- `gpurun --gpus 2` runs it."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _same_bits(a, b):
    assert np.array_equal(a["pose"], b["pose"]) and np.array_equal(a["cov"], b["cov"])
    for key in ("quality", "n_iterations", "termination_reason", "n_pairings", "cov_singular"):
        assert a[key] == b[key], key


def test_world_size_one_equals_plain_align(icp, capi, oracle):
    import torch
    from mola_fe_lidar_b200 import scene
    A, B, _ = scene.make_pair_c1(seed=31, n=30000, sigma=0.01)
    comm = capi.Comm(icp, capi.comm_unique_id(), 1, 0)
    smap = capi.NativeShardedMap(comm, A, cell=4.0, interleaved=True)
    assert smap.local_size() == len(A)
    g_a, g_b = icp.upload(A), icp.upload(B)
    r_s = smap.align(g_b, np.zeros(6))
    r_p = icp.align(g_a, g_b, np.zeros(6))
    _same_bits(r_s, r_p)
    o = oracle.icp_align(oracle.Cloud(A), oracle.Cloud(B), np.zeros(6), oracle.default_params(), kdtree=True)
    assert np.abs(r_s["pose"][:3] - o["pose"][:3]).max() < 1e-5 and np.abs(r_s["pose"][3:] - o["pose"][3:]).max() < 1e-6
    assert r_s["n_iterations"] == o["n_iterations"] and r_s["quality"] == o["quality"]
    dev = torch.device("cuda", 0)
    for k in (1, 6):
        ks = torch.empty((len(B), k), dtype=torch.int64, device=dev)
        kp = torch.empty((len(B), k), dtype=torch.int64, device=dev)
        smap.knn_keys(g_b, k, 0.7, ks.data_ptr())
        icp.knn_keys_device(g_a, g_b, k, 0.7, kp.data_ptr())
        torch.cuda.synchronize()
        assert torch.equal(ks, kp)
    smap.close(), comm.close(), g_a.free(), g_b.free()


WORKER = r"""
import json, os, sys
import numpy as np
sys.path.insert(0, %(root)r)
import torch
import torch.distributed as dist
from mola_fe_lidar_b200 import capi, scene
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("gloo")          # only to hand the communicator id around; the data path is the library's NCCL
icp = capi.ICP(capi.default_params(), device=rank)
ids = [capi.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
comm = capi.Comm(icp, ids[0], world, rank)
scans, poses = scene.make_sequence(4, seed=1)
themap = np.concatenate([(s.astype(np.float64) @ T[:3, :3].T + T[:3, 3]).astype(np.float32) for s, T in zip(scans[:3], poses[:3])])
local = scans[3]
truth = scene.matrix_to_pose6(poses[3])
guess = truth + np.array([0.1, -0.05, 0.02, 0.003, 0, 0])
out = {"rank": rank}
for interleaved in (True, False):
    smap = capi.NativeShardedMap(comm, themap, cell=4.0, interleaved=interleaved)
    out["shard_%%d" %% int(interleaved)] = smap.local_size()
    g_map, g_loc = icp.upload(themap), icp.upload(local)
    r_s = smap.align(g_loc, guess)
    r_p = icp.align(g_map, g_loc, guess)
    same = bool(np.array_equal(r_s["pose"], r_p["pose"]) and np.array_equal(r_s["cov"], r_p["cov"]) and
                all(r_s[k] == r_p[k] for k in ("quality", "n_iterations", "termination_reason", "n_pairings")))
    out["align_same_%%d" %% int(interleaved)] = same
    out["iters"] = int(r_s["n_iterations"]); out["err_m"] = float(np.abs(r_s["pose"][:3] - truth[:3]).max())
    for k in (1, 6):
        ks = torch.empty((len(local), k), dtype=torch.int64, device=dev)
        kp = torch.empty((len(local), k), dtype=torch.int64, device=dev)
        smap.knn_keys(g_loc, k, 0.7, ks.data_ptr(), pose6=guess)
        icp.knn_keys_device(g_map, g_loc, k, 0.7, kp.data_ptr(), pose6=guess)
        torch.cuda.synchronize()
        out["keys_same_%%d_%%d" %% (int(interleaved), k)] = bool(torch.equal(ks, kp))
    smap.close(); g_map.free(); g_loc.free()
comm.close(); icp.close()
res = [None] * world
dist.all_gather_object(res, out)
if rank == 0:
    print("WORKER " + json.dumps(res))
dist.destroy_process_group()
"""


def test_two_ranks_bit_identical_to_unsharded(capi, tmp_path):
    if capi.device_count() < 2:
        pytest.skip("needs two GPUs: NCCL does not place two ranks on one device (run under `gpurun --gpus 2`)")
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, (p.stdout[-2000:], p.stderr[-3000:])
    res = json.loads([l for l in p.stdout.splitlines() if l.startswith("WORKER ")][-1][len("WORKER "):])
    assert len(res) == 2
    for r in res:
        for key, val in r.items():
            if key.startswith("align_same") or key.startswith("keys_same"):
                assert val, (key, r)
        assert r["err_m"] < 0.05 and r["iters"] >= 1
    # a real split: both ranks hold a part of the map
    assert all(r["shard_1"] > 0 and r["shard_0"] > 0 for r in res)
