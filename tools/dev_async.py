"""Developer probe: module throughput with the asynchronous feed, upload prefetch on / off."""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mola_fe_lidar_b200 import capi, lidar_odometry, scene
N = 44
scans, poses = scene.make_sequence(N, seed=1)
h = [torch.from_numpy(np.ascontiguousarray(s.T)).pin_memory() for s in scans]
n = len(scans[0])
for mode in ("async+prefetch", "async", "sync"):
    extra = "  b200_extra_edge_checks: false\n"
    if mode == "async":
        extra += "  b200_prefetch_uploads: false\n"
    lo = lidar_odometry.LidarOdometry(yaml_text=lidar_odometry.system_yaml(extra=extra))
    def feed(i):
        a = (h[i][0].data_ptr(), h[i][1].data_ptr(), h[i][2].data_ptr(), n, 0.1 * i)
        if mode == "sync":
            lo.onNewObservationSoA(*a, sync=True)
        else:
            lo.enqueueObservationSoA(*a)
    for i in range(12):
        feed(i)
    lo.wait_idle()
    torch.cuda.synchronize()
    t = time.time()
    for i in range(12, N):
        feed(i)
    lo.wait_idle()
    dt = time.time() - t
    st = lo.state()
    pr = lo.profile()
    print("%-15s %.3f ms/scan  (%d processed, %d icp)" % (mode, dt / (N - 12) * 1e3, st["n_processed"], st["n_icp"]),
          {k.replace("doProcessNewObservation.", ""): (round(v[1] / max(v[0], 1) * 1e3, 3), round(v[2] * 1e3, 2)) for k, v in pr.items() if v[0] > 0 and v[1] > 1e-4})
    lo.close()
