"""Wall time of one registration (velocity-like guess) -- probe for the fit / search overlap (B200ICP_OVERLAP)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mola_fe_lidar_b200 import capi, scene
icp = capi.ICP(capi.default_params())
scans, poses = scene.make_sequence(3, seed=1)
a, b = icp.upload(scans[1]), icp.upload(scans[2])
T = np.linalg.inv(poses[1]) @ poses[2]
guess = scene.matrix_to_pose6(T @ scene.pose_matrix(0.03, -0.01, 0.005, np.deg2rad(0.05)))
ts = []
for rep in range(60):
    t = time.perf_counter(); r = icp.align(a, b, guess); ts.append((time.perf_counter() - t) * 1e3)
ts = np.array(ts[10:])
print("overlap=%s graph=%s: align wall ms median %.3f min %.3f  iters %d pose %s" % (
    os.environ.get("B200ICP_OVERLAP", "0"), os.environ.get("B200ICP_GRAPH", "1"), np.median(ts), ts.min(), r["n_iterations"],
    np.round(r["pose"][:3], 6)))
