"""Writes the golden fixtures of tests/golden/ (committed next to this script).

knn_c1.npz : radius-capped 6-NN of 300 queries in a 5,000-point sub-scan, from
             scipy.spatial.cKDTree in float64, kept only where the neighbour
             order is unambiguous in float32 (gaps > 1e-4 m), with d2 re-evaluated
             in the float32 order of SURVEY Appendix A.3 by plain numpy.
icp_c1.json: the oracle's own registration of a C1 pair, frozen after it was
             checked against the analytic transform (regression pin: the
             reference has no golden vectors to pin against).
"""
import json
import os
import sys

import numpy as np
from scipy.spatial import cKDTree

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle_api as O  # noqa: E402
from mola_fe_lidar_b200 import scene  # noqa: E402

A, B, pose = scene.make_pair_c1(seed=7, n=5000, sigma=0.01)
R, t = O.pose_to_Rt(pose)
q_all = O.transform_points(R, t, B)
tree = cKDTree(A.astype(np.float64))
dd, ii = tree.query(q_all.astype(np.float64), k=7, distance_upper_bound=0.8)
ok = np.isfinite(dd[:, :7]).all(axis=1) & (np.diff(dd, axis=1).min(axis=1) > 1e-4) & (dd[:, 5] < 0.69) & (dd[:, 6] > 0.71)
ok |= np.isfinite(dd[:, :6]).all(axis=1) & (np.diff(dd[:, :7], axis=1).min(axis=1) > 1e-4) & (dd[:, 5] < 0.69)
sel = np.where(ok)[0][:300]
qry = q_all[sel]
idx = ii[sel, :6].astype(np.uint32)
idx[dd[sel, :6] > 0.7] = 0xFFFFFFFF
p = A[np.minimum(idx, len(A) - 1)]
d = qry[:, None, :] - p
d2 = ((d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]).astype(np.float32) + d[..., 2] * d[..., 2]).astype(np.float32)
d2[idx == 0xFFFFFFFF] = np.inf
np.savez_compressed(os.path.join(HERE, "knn_c1.npz"), ref=A, qry=qry, idx=idx, d2_bits=d2.view(np.uint32))

meta = dict(seed=3, n=4000, sigma=0.0)
A, B, pose = scene.make_pair_c1(**meta)
r = O.icp_align(O.Cloud(A), O.Cloud(B), np.zeros(6), O.default_params(), kdtree=True)
assert np.abs(r["pose"][:3] - pose[:3]).max() < 5e-3
meta.update(pose=[float(v) for v in r["pose"]], n_iterations=int(r["n_iterations"]),
            n_pairings=int(r["n_pairings"]), quality=float(r["quality"]), truth=[float(v) for v in pose])
json.dump(meta, open(os.path.join(HERE, "icp_c1.json"), "w"), indent=1)
print("golden written:", len(sel), "knn queries;", meta)

# edges_planes_c1.npz : FilterEdgesPlanes classes of a 20,000-point sub-scan from an INDEPENDENT evaluation (numpy
#             eigvalsh / eigh on the per-voxel covariance in float64, voxels found with np.unique), kept only where
#             every gate is away from its threshold by more than 2 % (so that the Jacobi solver of the oracle / the
#             device and LAPACK must agree): per voxel the class (0 none, 1 edges, 2 planes), per point the voxel.
A = scene.make_pair_c1(seed=5, n=20000, sigma=0.01)[0]
res, mx20, mx10, mn20, mn10, minpts = np.float32(1.0), 30.0, 30.0, 80.0, 80.0, 5
vox = np.floor(A / res).astype(np.int64)
uniq, inv, cnt = np.unique(vox, axis=0, return_inverse=True, return_counts=True)
inv = inv.reshape(-1)
cls = np.zeros(len(uniq), dtype=np.int8)
sure = np.zeros(len(uniq), dtype=bool)
for v in range(len(uniq)):
    if cnt[v] < minpts:
        sure[v] = True
        continue
    P = A[inv == v].astype(np.float64)
    C = np.cov(P.T, bias=True)
    w, V = np.linalg.eigh(C)
    e0, e1, e2 = w
    if e0 <= 0:
        continue  # degenerate: leave it to the parity tests
    m = 0.02
    is_edge = e2 < mx20 * e0 and e1 < mx10 * e0
    is_plane = (not is_edge) and e2 > mn20 * e0 and e1 > mn10 * e0 and abs(V[2, 0]) < 0.9
    margins = [abs(e2 / (mx20 * e0) - 1), abs(e1 / (mx10 * e0) - 1), abs(e2 / (mn20 * e0) - 1), abs(e1 / (mn10 * e0) - 1),
               abs(abs(V[2, 0]) / 0.9 - 1)]
    sure[v] = min(margins) > m
    cls[v] = 1 if is_edge else (2 if is_plane else 0)
np.savez_compressed(os.path.join(HERE, "edges_planes_c1.npz"), pts=A, voxel_of_point=inv.astype(np.int32),
                    voxel_class=cls, voxel_sure=sure)
print("golden written: edges/planes,", int(sure.sum()), "of", len(uniq), "voxels unambiguous;",
      int((cls[sure] == 1).sum()), "edges,", int((cls[sure] == 2).sum()), "planes")
