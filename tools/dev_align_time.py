"""Developer probe: align / search timings on 3 consecutive 120k scans (one process per variant)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mola_fe_lidar_b200 import capi, scene

def main():
    icp = capi.ICP(capi.default_params())
    scans, poses = scene.make_sequence(4, seed=1)
    rad = float(os.environ.get('B200ICP_DEV_RADIUS', '0'))
    clouds = [icp.upload(s, search_radius=rad) for s in scans]
    icp.profile_enable(True)
    tot = []
    for i in range(1, 4):
        for rep in range(3):
            icp.profile_reset()
            r = icp.align(clouds[i - 1], clouds[i], np.zeros(6))
        p = icp.profile()
        tot.append("%.3f/%.3f/%.3f" % (p["match_ms"] / max(p["match_launches"], 1), p["fit_ms"] / max(p["fit_launches"], 1), p["solve_ms"] / max(p["solve_launches"], 1)))
    icp.profile_enable(False)
    ts = []
    for rep in range(5):
        t = time.time(); r = icp.align(clouds[0], clouds[1], np.zeros(6)); ts.append((time.time() - t) * 1e3)
    print(os.environ.get("B200ICP_WALK", "-"), "radius", rad, "search/fit/solve ms:", " ".join(tot), "| align wall min %.2f ms iters %d" % (min(ts), r["n_iterations"]))
main()
