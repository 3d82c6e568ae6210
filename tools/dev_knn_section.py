"""Developer probe: bench.py's kNN / scan-to-map section alone."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from mola_fe_lidar_b200 import capi
bench.N_SCANS = 10
scans, poses = bench.make_scans(1)
icp = capi.ICP(capi.default_params())
knn, c3 = bench.knn_microbench(torch, icp, scans, poses, torch.device("cuda", 0), 6547.8)
for k in knn:
    print(json.dumps(k))
print(json.dumps(c3))
