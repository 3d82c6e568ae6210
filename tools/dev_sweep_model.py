"""CPU model of the seeded item sweep (search v3): for consecutive 120k-pt scans, how many points would one warp
stage per item when it sweeps the cells of its queries' box grown by a margin, and how many lanes such a sweep covers.
Development aid only (numpy / scipy); prints distributions used in DESIGN.md section 4.1."""
import importlib
import sys

import numpy as np
from scipy.spatial import cKDTree

sys.path.insert(0, ".")
scene = importlib.import_module("mola_fe_lidar_b200.scene")

R = 0.70
BLOCK = R * 1.002
CELL = BLOCK / 4


def spread10(v):
    v = v & 0x3FF
    v = (v | (v << 16)) & 0x030000FF
    v = (v | (v << 8)) & 0x0300F00F
    v = (v | (v << 4)) & 0x030C30C3
    v = (v | (v << 2)) & 0x09249249
    return v


def own_order(P):
    o = P.min(axis=0)
    h = np.floor((P - o) / CELL * 2).astype(np.int64)
    f = h >> 1
    mort = spread10(f[:, 0] >> 2) | (spread10(f[:, 1] >> 2) << 1) | (spread10(f[:, 2] >> 2) << 2)
    sub = (f[:, 0] & 3) | ((f[:, 1] & 3) << 2) | ((f[:, 2] & 3) << 4)
    octant = (h[:, 0] & 1) | ((h[:, 1] & 1) << 1) | ((h[:, 2] & 1) << 2)
    key = (((mort << 6) | sub) << 3) | octant
    return np.argsort(key, kind="stable")


def main():
    scans, poses = scene.make_sequence(4, seed=1)
    G = scans[1].astype(np.float64)
    L = scans[2].astype(np.float64)
    T = np.linalg.inv(poses[1]) @ poses[2]
    # guess error of a velocity model / an early iteration: 3 cm, 0.05 deg
    Tg = T @ scene.pose_matrix(0.03, -0.01, 0.005, np.deg2rad(0.05))
    for label, Tq in (("guess (3 cm off)", Tg), ("converged", T)):
        order = own_order(L)
        Q = (L @ Tq[:3, :3].T + Tq[:3, 3])[order]
        tree = cKDTree(G)
        d, _ = tree.query(Q, k=6, distance_upper_bound=R)
        r6 = np.where(np.isfinite(d[:, 5]), d[:, 5], R)
        og = G.min(axis=0)
        gc = np.floor((G - og) / CELL).astype(np.int64)
        cells, counts = np.unique(gc, axis=0, return_counts=True)
        qc = np.floor((Q - og) / CELL).astype(np.int64)
        qf = (Q - og) / CELL - qc
        n_items = len(Q) // 32
        for mode in ("seeded", "round0"):
            staged = np.zeros(n_items, dtype=np.int64)
            uncovered = np.zeros(n_items, dtype=np.int64)
            margin = np.zeros(n_items, dtype=np.int64)
            for it in range(n_items):
                s = slice(32 * it, 32 * it + 32)
                lo, hi = qc[s].min(axis=0), qc[s].max(axis=0)
                if mode == "seeded":
                    cap = np.minimum(r6[s] + 0.02, R)
                else:
                    # round 0: k-th best among the points of the box's own cells
                    inbox = np.all((cells >= lo) & (cells <= hi), axis=1)
                    if counts[inbox].sum() >= 6:
                        sel = np.all((gc >= lo) & (gc <= hi), axis=1)
                        pts = G[sel]
                        dd = np.sqrt(((Q[s][:, None, :] - pts[None, :, :]) ** 2).sum(-1))
                        dd.sort(axis=1)
                        cap = np.minimum(dd[:, 5], R)
                    else:
                        cap = np.full(32, R)
                capc = cap / CELL
                # margin each lane needs so that its ball lies inside the box of cells [lo - S, hi + S]
                need = np.zeros(32)
                for ax in range(3):
                    below = capc - ((qc[s][:, ax] - lo[ax]) + qf[s][:, ax])
                    above = capc - ((hi[ax] - qc[s][:, ax]) + 1 - qf[s][:, ax])
                    need = np.maximum(need, np.maximum(below, above))
                needS = np.ceil(np.maximum(need, 0)).astype(np.int64)
                # largest margin whose staged count fits the budget
                best = None
                for S in sorted(set(needS.tolist()), reverse=True):
                    inbox = np.all((cells >= lo - S) & (cells <= hi + S), axis=1)
                    c = counts[inbox].sum()
                    if c <= 768 or S == 0:
                        best = (S, c)
                        break
                if best is None:
                    best = (0, 0)
                margin[it], staged[it] = best
                uncovered[it] = (needS > best[0]).sum()
            pct = lambda a: np.percentile(a, [50, 90, 99, 100]).astype(int).tolist()
            print(f"{label:18s} {mode:7s} staged/item mean {staged.mean():.0f} p50/90/99/max {pct(staged)}  "
                  f"margin cells mean {margin.mean():.2f}  uncovered lanes total {uncovered.sum()} "
                  f"({100.0 * uncovered.sum() / len(Q):.2f} %), items with any {np.count_nonzero(uncovered)}")


if __name__ == "__main__":
    main()
