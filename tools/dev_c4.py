"""A slice of BASELINE config C4 (loop-closure candidates x Monte-Carlo guesses in one b200icp_align_batch) --
the command profiled under ncu for the batched regime."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mola_fe_lidar_b200 import capi, lidar_odometry, scene  # noqa: E402

n_pairs, mc = int(os.environ.get("PAIRS", "64")), 10
icp = capi.ICP(yaml_text=open(os.path.join(lidar_odometry.PARAMS_DIR, "icp-settings-loop-closure.yaml")).read())
scans, _ = scene.make_sequence(8, seed=1)
rng = np.random.default_rng(99)
sub = [s[np.sort(rng.choice(len(s), size=20000, replace=False))] for s in scans]
clouds = [icp.upload(x) for x in sub]
g = np.zeros((n_pairs * mc, 6))
g[:, :3] = rng.normal(0.0, 3.0, size=(n_pairs * mc, 3))
g[:, 3] = rng.normal(0.0, np.deg2rad(2.0), size=n_pairs * mc)
g[:, 0] += 1.0
fr, to = [], []
for p in range(n_pairs):
    i = int(rng.integers(0, len(sub) - 1))
    fr += [clouds[i]] * mc
    to += [clouds[i + 1]] * mc
icp.align_batch(fr, to, g)
icp.profile_enable(True)
icp.profile_reset()
t = time.time()
res = icp.align_batch(fr, to, g)
dt = time.time() - t
p = icp.profile()
its = sum(r["n_iterations"] + 1 for r in res)
print("%d registrations in %.1f ms = %.0f /s; %d outer iterations; search %.1f ms fit %.1f ms over %d / %d launches" % (
    len(res), dt * 1e3, len(res) / dt, its, p["match_ms"], p["fit_ms"], p["match_launches"], p["fit_launches"]))
