"""Developer probe: where the module's per-scan time goes (host clock per section)."""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mola_fe_lidar_b200 import capi, lidar_odometry, scene
scans, poses = scene.make_sequence(6, seed=1)
h = [torch.from_numpy(np.ascontiguousarray(s.T)).pin_memory() for s in scans]
n = len(scans[0])
for extra in ("  b200_extra_edge_checks: false\n", ""):
    lo = lidar_odometry.LidarOdometry(yaml_text=lidar_odometry.system_yaml(extra=extra))
    ts = []
    for s in range(40):
        i = s % 10; i = i if i < 6 else 10 - i
        t = time.time()
        lo.onNewObservationSoA(h[i][0].data_ptr(), h[i][1].data_ptr(), h[i][2].data_ptr(), n, 0.1 * s, sync=True)
        ts.append((time.time() - t) * 1e3)
    lo.wait_idle()
    print("extra=%r per-scan wall ms:" % extra, np.round(ts, 2).tolist())
    pr = lo.profile()
    print({k: (v[0], round(v[1] / max(v[0], 1) * 1e3, 3)) for k, v in pr.items()})
    lo.close()
# raw C ABI from pinned host memory
icp = capi.ICP(capi.default_params())
prev = icp.upload_ptrs(h[0][0].data_ptr(), h[0][1].data_ptr(), h[0][2].data_ptr(), n)
ts, tu = [], []
for s in range(1, 20):
    i = s % 10; i = i if i < 6 else 10 - i
    t = time.time()
    cur = icp.upload_ptrs(h[i][0].data_ptr(), h[i][1].data_ptr(), h[i][2].data_ptr(), n)
    t1 = time.time()
    r = icp.align(prev, cur, np.zeros(6))
    t2 = time.time()
    prev.free(); prev = cur
    tu.append((t1 - t) * 1e3); ts.append((t2 - t1) * 1e3)
print("C ABI upload ms:", np.round(tu, 2).tolist())
print("C ABI align  ms:", np.round(ts, 2).tolist())
