"""An INDEPENDENT re-implementation of the whole registration -- written from
SURVEY.md Appendix A (A.2-A.8) in numpy / scipy, sharing no code with
oracle/icp_oracle.c -- run next to the C oracle on the same inputs.

The reference's own arithmetic (mp2p_icp / MRPT) is absent from this container
and the reference ships no golden vectors, so the oracle cannot be pinned
against it.  What can be done is to restate the frozen
algorithm twice, in two languages with different building blocks (cKDTree +
float32 re-ranking instead of the oracle's kd-tree, LAPACK `eigh` instead of
Jacobi, `lstsq` instead of Householder QR, `scipy.linalg.expm / logm` instead
of closed-form SE(3) maps), and require both to agree: same pairing decisions,
same iteration count and termination, same quality, poses within 1e-7."""
import numpy as np
import pytest
from scipy.linalg import expm, logm
from scipy.spatial import cKDTree


def _rot_ypr(yaw, pitch, roll):  # A.2: R = Rz(yaw) Ry(pitch) Rx(roll)
    cy, sy, cp, sp, cr, sr = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch), np.cos(roll), np.sin(roll)
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1.0]])
    Ry = np.array([[cp, 0, sp], [0, 1.0, 0], [-sp, 0, cp]])
    Rx = np.array([[1.0, 0, 0], [0, cr, -sr], [0, sr, cr]])
    return Rz @ Ry @ Rx


def _hat(w):
    return np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0.0]])


def _se3_exp(eps):  # eps = (v, w)
    X = np.zeros((4, 4))
    X[:3, :3] = _hat(eps[3:])
    X[:3, 3] = eps[:3]
    return expm(X)


def _se3_log(T):
    X = np.real(logm(T))
    return np.array([X[0, 3], X[1, 3], X[2, 3], X[2, 1], X[0, 2], X[1, 0]])


def _knn_f32(tree, G32, q32, k, cap_d2):
    """A.3/A.4: the k smallest (d2 in float32 with the fixed operation order,
    index) keys with d2 <= cap.  Candidates come from a double-precision tree
    with a safety margin, then are re-ranked exactly."""
    kk = min(k + 6, len(G32))
    _, cand = tree.query(q32.astype(np.float64), k=kk, distance_upper_bound=np.sqrt(float(cap_d2)) * 1.01)
    out = []
    for i in range(len(q32)):
        c = cand[i][cand[i] < len(G32)]
        d = q32[i] - G32[c]                                  # float32 differences
        d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]  # float32, ((dx2 + dy2) + dz2)
        order = np.lexsort((c, d2))
        keep = [(d2[j], c[j]) for j in order if d2[j] <= cap_d2][:k]
        out.append(keep)
    return out


def _match_point2plane(tree, G32, L32, R, t, thr, eig_thr, knn, min_pts):
    """A.5.  Returns (indices of paired local points, centroids, normals)."""
    q = (L32.astype(np.float64) @ R.T + t).astype(np.float32)   # A.2: one rounding to float32
    thr2 = np.float32(thr) * np.float32(thr)
    nb = _knn_f32(tree, G32, q, knn, thr2)
    idx, cen, nor = [], [], []
    for i, lst in enumerate(nb):
        if len(lst) < min_pts:
            continue
        P = G32[[j for _, j in lst]].astype(np.float64)
        c = P.mean(axis=0)
        C = (P - c).T @ (P - c) / len(P)
        w, V = np.linalg.eigh(C)                              # ascending
        if w[0] > eig_thr * w[2]:
            continue
        n = V[:, 0]
        if abs(n @ (q[i].astype(np.float64) - c)) > thr:
            continue
        idx.append(i), cen.append(c), nor.append(n)
    return np.array(idx, dtype=int), np.array(cen).reshape(-1, 3), np.array(nor).reshape(-1, 3)


def _gauss_newton(P, Cc, Nn, T, max_iters=20, min_delta=1e-10):
    """A.6: right perturbation T (+) exp(eps), eps = (v, w)."""
    for _ in range(max_iters):
        R, t = T[:3, :3], T[:3, 3]
        r = np.einsum("ij,ij->i", Nn, P @ R.T + t - Cc)
        nR = Nn @ R                                           # rows n^T R
        J = np.hstack([nR, -np.cross(nR, P)])                 # d r / d v = n^T R ; d r / d w = -n^T R [p]x
        delta = np.linalg.lstsq(J.T @ J, -(J.T @ r), rcond=None)[0]
        T = T @ _se3_exp(delta)
        if np.linalg.norm(delta) < min_delta:
            break
    return T


def independent_icp(G32, L32, guess, thr=0.70, eig_thr=0.07, knn=6, min_pts=3, max_it=100, step_t=5e-5,
                    step_r=1e-5, q_thr=0.10):
    tree = cKDTree(G32.astype(np.float64))
    T = np.eye(4)
    T[:3, :3] = _rot_ypr(*guess[3:])
    T[:3, 3] = guess[:3]
    it, reason, first_pairs = 0, 3, None
    while it < max_it:
        idx, cen, nor = _match_point2plane(tree, G32, L32, T[:3, :3], T[:3, 3], thr, eig_thr, knn, min_pts)
        if first_pairs is None:
            first_pairs = idx
        if len(idx) == 0:
            reason = 1
            break
        Tn = _gauss_newton(L32[idx].astype(np.float64), cen, nor, T.copy())
        d = _se3_log(np.linalg.inv(T) @ Tn)
        T = Tn
        if np.linalg.norm(d[:3]) < step_t and np.linalg.norm(d[3:]) < step_r:
            reason = 4                                         # Stalled: the iteration counter is not advanced
            break
        it += 1
    # A.8: fraction of local points with a neighbour at d2 < thr^2 (strict)
    q = (L32.astype(np.float64) @ T[:3, :3].T + T[:3, 3]).astype(np.float32)
    q2 = np.float32(q_thr) * np.float32(q_thr)
    hits = sum(1 for lst in _knn_f32(tree, G32, q, 1, q2) if lst and lst[0][0] < q2)
    return T, it, reason, hits / len(L32), first_pairs


def _scene(rng, n=2500):
    """Three walls and a floor with a few boxes: planar enough for the matcher."""
    pts = []
    for _ in range(n):
        s = rng.integers(0, 4)
        u, v = rng.uniform(-6, 6), rng.uniform(-6, 6)
        p = [(u, v, -1.5), (6.0, u, v / 3 + 0.5), (u, -6.0, v / 3 + 0.5), (-6.0 + 0.1 * v, u, v / 3 + 0.5)][s]
        pts.append(p)
    return (np.array(pts) + rng.normal(0, 0.01, size=(n, 3))).astype(np.float32)


@pytest.mark.parametrize("seed,guess", [(1, np.zeros(6)), (2, np.array([0.05, -0.03, 0.0, 0.004, 0.0, 0.0]))])
def test_oracle_agrees_with_independent_numpy_icp(oracle, seed, guess):
    rng = np.random.default_rng(seed)
    G = _scene(rng)
    truth = np.array([0.12, -0.08, 0.03, np.deg2rad(0.8), np.deg2rad(0.2), np.deg2rad(-0.15)])
    Rt, tt = _rot_ypr(*truth[3:]), truth[:3]
    sel = rng.choice(len(G), size=1800, replace=False)
    L = ((G[sel].astype(np.float64) - tt) @ Rt + rng.normal(0, 0.005, size=(1800, 3))).astype(np.float32)
    # 10 outer iterations at most: enough to converge or to enter the 2-cycle between two pairing sets that
    # ICP sometimes settles in (then both implementations must report MaxIterations after the same steps)
    prm = oracle.default_params(max_iterations=10)
    go, lo = oracle.Cloud(G), oracle.Cloud(L)
    ro = oracle.icp_align(go, lo, guess, prm, kdtree=True)
    T, it, reason, quality, first_pairs = independent_icp(G, L, guess, max_it=10)
    # pairing decisions of the first matcher run
    R0, t0 = oracle.pose_to_Rt(guess)
    mo = oracle.match_point2plane(go, lo, R0, t0, prm, kdtree=True)
    assert len(first_pairs) > 1500
    assert np.array_equal(np.nonzero(mo["paired"])[0], first_pairs)
    # the whole registration: same path, step for step
    assert ro["termination_reason"] == reason and reason in (3, 4)
    assert ro["n_iterations"] == it
    assert ro["quality"] == pytest.approx(quality, abs=1e-12)
    assert np.abs(ro["t"] - T[:3, 3]).max() < 1e-7
    assert np.abs(ro["R"] - T[:3, :3]).max() < 1e-7
    assert np.abs(ro["pose"][:3] - truth[:3]).max() < 5e-3   # and both are right


# ---------------------------------------------------------------------------
# rows M, N: Matcher_Points_DistanceThreshold + Solver_Horn with the scale-outlier
# rule, restated independently (the rotation comes from an SVD / Kabsch solve
# instead of Horn's quaternion eigenvector: the same optimum by a different road)
def _match_points(tree, G32, L32, R, t, thr):
    q = (L32.astype(np.float64) @ R.T + t).astype(np.float32)
    thr2 = np.float32(thr) * np.float32(thr)
    nb = _knn_f32(tree, G32, q, 1, thr2)
    idx = [i for i, lst in enumerate(nb) if lst and lst[0][0] < thr2]        # strict (A.8 / orc_match_points)
    nn = [nb[i][0][1] for i in idx]
    return np.array(idx, dtype=int), np.array(nn, dtype=int)


def _horn_kabsch(P, Q, scale_thr=1.1):
    """A.10: centroids over ALL pairs; a pair whose centroid-relative norms
    differ by more than the threshold ratio is left out of the rotation."""
    pc, qc = P.mean(axis=0), Q.mean(axis=0)
    b, a = P - pc, Q - qc
    bn, an = np.linalg.norm(b, axis=1), np.linalg.norm(a, axis=1)
    mx, mn = np.maximum(bn, an), np.minimum(bn, an)
    keep = (mn > 0) & (mx / np.where(mn > 0, mn, 1.0) <= scale_thr)
    if keep.sum() < 3:
        return None
    S = b[keep].T @ a[keep]                                      # sum b a^T
    U, _, Vt = np.linalg.svd(S)
    D = np.diag([1.0, 1.0, np.sign(np.linalg.det(Vt.T @ U.T))])
    R = Vt.T @ D @ U.T
    T = np.eye(4)
    T[:3, :3], T[:3, 3] = R, qc - R @ pc
    return T


def independent_icp_p2p_horn(G32, L32, guess, thr, max_it, step_t=5e-5, step_r=1e-5, q_thr=0.10):
    tree = cKDTree(G32.astype(np.float64))
    T = np.eye(4)
    T[:3, :3] = _rot_ypr(*guess[3:])
    T[:3, 3] = guess[:3]
    it, reason = 0, 3
    while it < max_it:
        idx, nn = _match_points(tree, G32, L32, T[:3, :3], T[:3, 3], thr)
        if len(idx) == 0:
            reason = 1
            break
        Tn = _horn_kabsch(L32[idx].astype(np.float64), G32[nn].astype(np.float64))
        if Tn is None:
            reason = 2
            break
        d = _se3_log(np.linalg.inv(T) @ Tn)
        T = Tn
        if np.linalg.norm(d[:3]) < step_t and np.linalg.norm(d[3:]) < step_r:
            reason = 4
            break
        it += 1
    q = (L32.astype(np.float64) @ T[:3, :3].T + T[:3, 3]).astype(np.float32)
    q2 = np.float32(q_thr) * np.float32(q_thr)
    hits = sum(1 for lst in _knn_f32(tree, G32, q, 1, q2) if lst and lst[0][0] < q2)
    return T, it, reason, hits / len(L32)


def test_oracle_p2p_horn_agrees_with_independent_kabsch_icp(oracle):
    rng = np.random.default_rng(5)
    G = _scene(rng, 2200)
    truth = np.array([0.10, 0.06, -0.02, np.deg2rad(0.6), np.deg2rad(-0.2), np.deg2rad(0.1)])
    Rt, tt = _rot_ypr(*truth[3:]), truth[:3]
    sel = rng.choice(len(G), size=1500, replace=False)
    L = ((G[sel].astype(np.float64) - tt) @ Rt + rng.normal(0, 0.004, size=(1500, 3))).astype(np.float32)
    prm = oracle.default_params(max_iterations=8, solver_kind=1, matcher_kind=1, distance_threshold=0.5)
    ro = oracle.icp_align(oracle.Cloud(G), oracle.Cloud(L), np.zeros(6), prm, kdtree=True)
    T, it, reason, quality = independent_icp_p2p_horn(G, L, np.zeros(6), 0.5, 8)
    assert ro["termination_reason"] == reason and reason in (3, 4)
    assert ro["n_iterations"] == it
    assert ro["quality"] == pytest.approx(quality, abs=1e-12)
    assert np.abs(ro["t"] - T[:3, 3]).max() < 1e-7
    assert np.abs(ro["R"] - T[:3, :3]).max() < 1e-7
    assert np.abs(ro["pose"][:3] - truth[:3]).max() < 2e-2
