"""One registration of two consecutive 120k-pt scans from a constant-velocity-like guess (3 cm / 0.05 deg off the
truth, as in bench.py's value leg) -- the command profiled under ncu."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mola_fe_lidar_b200 import capi, scene  # noqa: E402

icp = capi.ICP(capi.default_params())
scans, poses = scene.make_sequence(3, seed=1)
a, b = icp.upload(scans[1]), icp.upload(scans[2])
T = np.linalg.inv(poses[1]) @ poses[2]
guess = scene.matrix_to_pose6(T @ scene.pose_matrix(0.03, -0.01, 0.005, np.deg2rad(0.05)))
prof = bool(os.environ.get("B200ICP_PROFILE"))
if prof:
    icp.align(a, b, guess)
    icp.profile_enable(True)
    icp.profile_reset()
for rep in range(int(os.environ.get("REPS", "1"))):
    r = icp.align(a, b, guess)
if prof:
    p = icp.profile()
    print("search/fit ms per launch: %.4f / %.4f over %d launches" % (p["match_ms"] / max(p["match_launches"], 1), p["fit_ms"] / max(p["fit_launches"], 1), p["match_launches"]))
print(r["n_iterations"], r["pose"], r["quality"])
if os.environ.get("KNN"):
    idx, d2 = icp.knn(a, b, 6, 0.7, pose6=guess)
