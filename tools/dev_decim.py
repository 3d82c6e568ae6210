"""Developer probe: the decimated regime (1.0 m voxels -> ~3k points per cloud) at the C ABI."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mola_fe_lidar_b200 import capi, scene
icp = capi.ICP(capi.default_params())
scans, poses = scene.make_sequence(16, seed=1)
res = float(os.environ.get("B200ICP_DEV_VOXEL", "1.0"))
raw = [icp.upload(s) for s in scans]
tv = []
dec = []
for r in raw:
    t = time.time(); d = icp.voxel_decimate(r, res); tv.append((time.time() - t) * 1e3); dec.append(d)
print("voxel %.2f m: points" % res, [len(d) for d in dec][:6], "voxel_decimate wall ms", np.round(tv, 3).tolist())
guess = np.zeros(6)
for prof in (False, True):
    icp.profile_enable(prof); icp.profile_reset()
    ta, its = [], []
    guess = np.zeros(6)
    for i in range(1, 16):
        t = time.time(); r = icp.align(dec[i - 1], dec[i], guess); ta.append((time.time() - t) * 1e3)
        its.append(r["n_iterations"] + 1)
        guess = np.array([r["pose"][0], r["pose"][1], r["pose"][2], r["pose"][3], 0, 0])
    print("profiling", prof, "align wall ms", np.round(ta, 3).tolist(), "matcher runs", its)
    if prof:
        p = icp.profile()
        print({k: round(v, 4) if isinstance(v, float) else v for k, v in p.items() if v})
        print("per launch: search %.4f fit+solve %.4f ms" % (p["match_ms"] / max(p["match_launches"], 1), p["fit_ms"] / max(p["fit_launches"], 1)))
