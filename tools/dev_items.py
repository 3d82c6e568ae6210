import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mola_fe_lidar_b200 import capi, scene
icp = capi.ICP(capi.default_params())
scans, poses = scene.make_sequence(3, seed=1)
clouds = [icp.upload(s) for s in scans]
icp.profile_enable(True)
for (a, b) in ((0, 1), (1, 2)):
    for k, r in ((6, 0.7), (1, 0.1)):
        icp.profile_reset()
        icp.knn(clouds[a], clouds[b], k, r)
        p = icp.profile()
        print("pair", a, b, "k", k, "r", r, "kernel ms %.3f" % p["knn_ms"], flush=True)
