"""One registration of two 120k-pt scans -- the command profiled under ncu."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mola_fe_lidar_b200 import capi, scene  # noqa: E402

icp = capi.ICP(capi.default_params())
scans, poses = scene.make_sequence(2, seed=1)
a, b = icp.upload(scans[0]), icp.upload(scans[1])
r = icp.align(a, b, np.zeros(6))
print(r["n_iterations"], r["pose"])
idx, d2 = icp.knn(a, b, 6, 0.7)
