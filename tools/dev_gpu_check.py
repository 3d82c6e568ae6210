"""Developer probe (not a test, not the bench): timings of the main calls on
120k-pt synthetic scans, with the library's own CUDA-event profile."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from mola_fe_lidar_b200 import capi, scene  # noqa: E402


def main():
    print("devices:", capi.device_count())
    icp = capi.ICP(capi.default_params())
    scans, poses = scene.make_sequence(4, seed=1)
    icp.profile_enable(True)
    t = time.time()
    clouds = [icp.upload(s) for s in scans]
    icp.synchronize()
    print("upload+index x4: %.2f ms each" % ((time.time() - t) * 1e3 / 4))
    for rep in range(3):
        t = time.time()
        c = icp.upload(scans[0])
        icp.synchronize()
        print("  upload+index: %.3f ms" % ((time.time() - t) * 1e3))
        c.free()
    for k, r in ((6, 0.7), (1, 0.7), (1, 0.1)):
        for rep in range(3):
            t = time.time()
            idx, d2 = icp.knn(clouds[0], clouds[1], k, r)
            dt = time.time() - t
        p = icp.profile()
        print("knn k=%d r=%.1f: wall %.2f ms, kernel avg %.3f ms" %
              (k, r, dt * 1e3, p["knn_ms"] / max(p["knn_launches"], 1)))
        icp.profile_reset()
    for i in range(1, 4):
        for rep in range(2):
            t = time.time()
            r = icp.align(clouds[i - 1], clouds[i], np.zeros(6))
            dt = time.time() - t
        p = icp.profile()
        gt = scene.relative_pose6(poses[i - 1], poses[i])
        print("align %d: wall %.2f ms iters=%d term=%s pairs=%d q=%.3f | search avg %.3f ms x%d, fit avg %.3f ms, solve avg %.3f ms | err %s" %
              (i, dt * 1e3, r["n_iterations"], capi.TERM[r["termination_reason"]], r["n_pairings"],
               r["quality"], p["match_ms"] / max(p["match_launches"], 1), p["match_launches"],
               p["fit_ms"] / max(p["fit_launches"], 1),
               p["solve_ms"] / max(p["solve_launches"], 1), np.round(r["pose"] - gt, 4)))
        icp.profile_reset()
    icp.profile_enable(False)
    for rep in range(3):
        t = time.time()
        r = icp.align(clouds[0], clouds[1], np.zeros(6))
        print("align (no profiling): wall %.2f ms iters=%d" % ((time.time() - t) * 1e3, r["n_iterations"]))
    # decimated variant
    t = time.time()
    dec = [icp.voxel_decimate(c, 1.0) for c in clouds]
    print("voxel 1.0 m: %d -> %d pts, %.2f ms each" % (len(clouds[0]), len(dec[0]), (time.time() - t) * 1e3 / 4))
    for rep in range(3):
        t = time.time()
        r = icp.align(dec[0], dec[1], np.zeros(6))
        print("align decimated: wall %.2f ms iters=%d q=%.3f" % ((time.time() - t) * 1e3, r["n_iterations"], r["quality"]))
    # batch of 64 small jobs
    A, B, pose = scene.make_pair_c1(seed=1, n=20000, sigma=0.01)
    ga, gb = icp.upload(A), icp.upload(B)
    rng = np.random.default_rng(0)
    for nb in (1, 16, 128):
        guesses = pose + np.c_[rng.normal(0, 0.1, (nb, 3)), rng.normal(0, 0.02, (nb, 1)), np.zeros((nb, 2))]
        for rep in range(2):
            t = time.time()
            rs = icp.align_batch([ga] * nb, [gb] * nb, guesses)
            dt = time.time() - t
        print("batch %d x 20k: wall %.2f ms (%.1f regs/s) iters max %d" %
              (nb, dt * 1e3, nb / dt, max(r["n_iterations"] for r in rs)))


if __name__ == "__main__":
    main()
