"""Two PROCESSES (torch.distributed.run, one rank each) over a map split by spatial cell: ShardedMap.query
(collective merge) and ShardedMap.query_fused (the search kernel stores into the peers' exchange buffers, CUDA
IPC) against the oracle's search of the UNSHARDED map, bit-exact.  With two GPUs the ranks
use NCCL on their own devices; on a one-GPU box both ranks share cuda:0 and the collectives go through gloo --
the kernels, the IPC mapping and the peer barrier are the same."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import json, os, sys
import numpy as np
sys.path.insert(0, %(root)r)
import torch
import torch.distributed as dist
from mola_fe_lidar_b200 import capi, multi_gpu as M
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
ndev = torch.cuda.device_count()
dev_id = rank %% ndev
torch.cuda.set_device(dev_id)
dev = torch.device("cuda", dev_id)
backend = "nccl" if ndev >= world else "gloo"
dist.init_process_group(backend)


class HostCollectives:
    # gloo moves host tensors: the same calls with a round trip through host memory (one-GPU boxes only)
    ReduceOp = dist.ReduceOp
    def all_reduce(self, t, op=None):
        c = t.cpu(); dist.all_reduce(c, op=op); t.copy_(c)
    def all_gather(self, lst, t):
        cl = [x.cpu() for x in lst]; dist.all_gather(cl, t.cpu())
        for a, b in zip(lst, cl): a.copy_(b)
    def all_gather_object(self, lst, obj): dist.all_gather_object(lst, obj)
    def barrier(self): dist.barrier()


coll = dist if backend == "nccl" else HostCollectives()
rng = np.random.default_rng(11)          # the same map and queries on every rank
themap = rng.uniform([-30, -30, -2], [30, 30, 2], size=(40000, 3)).astype(np.float32)
themap[200:220] = themap[100:120]        # exact duplicates: ties across shards
q = rng.uniform([-31, -31, -2], [31, 31, 2], size=(5000, 3)).astype(np.float32)
q[:10] = themap[200:210]
q[17] = np.nan
icp = capi.ICP(capi.default_params(), device=dev_id)
owner = M.partition_by_cell(themap, world, cell=5.0, mode="interleaved")
mine = M.shard_indices(owner, rank)
search = M.CudaShardSearch(icp, themap[mine], mine, 1.0, dev)
sm = M.ShardedMap(search, rank, world, coll)
qc = search.upload_queries(q, 1.0)
out = {"backend": backend, "shard": int(len(mine))}
for k in (1, 6):
    keys = sm.query(qc, k, 1.0)
    fused = sm.query_fused(qc, k, 1.0)
    torch.cuda.synchronize()
    out["same_%%d" %% k] = bool(torch.equal(keys, fused))
    if rank == 0:
        np.save(os.path.join(%(out)r, "keys_%%d.npy" %% k), keys.cpu().numpy())
dist.barrier()
sm.close(); qc.free(); search.close(); icp.close()
if rank == 0:
    np.save(os.path.join(%(out)r, "map.npy"), themap); np.save(os.path.join(%(out)r, "q.npy"), q)
    print("WORKER " + json.dumps(out))
dist.destroy_process_group()
"""


def test_two_process_sharded_map_matches_unsharded_oracle(oracle, tmp_path):
    from mola_fe_lidar_b200 import multi_gpu as M
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT, "out": str(tmp_path)})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29531", str(script)]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, (p.stdout[-2000:], p.stderr[-3000:])
    line = [l for l in p.stdout.splitlines() if l.startswith("WORKER ")][-1]
    info = json.loads(line[len("WORKER "):])
    assert info["same_1"] and info["same_6"], info          # fused peer-memory path == collective path
    themap, q = np.load(tmp_path / "map.npy"), np.load(tmp_path / "q.npy")
    for k in (1, 6):
        keys = np.load(tmp_path / f"keys_{k}.npy").view(np.uint64)
        gi, gd = M.unpack_keys(keys)
        idx, d2 = oracle.knn(oracle.Cloud(themap), q, k, np.float32(1.0), kdtree=True)
        assert np.array_equal(gi, idx)
        assert np.array_equal(gd.view(np.uint32), d2.view(np.uint32))
