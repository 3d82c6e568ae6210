"""CPU tests of the oracle (tests/ only): the restatement of SURVEY Appendix A
is cross-checked against independent implementations available here (scipy
cKDTree, numpy eigh / solve) and analytic known answers, because the reference
ships no golden vectors (parity unpinned)."""
import json
import os

import numpy as np
import pytest
from scipy.spatial import cKDTree

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_knn_kdtree_equals_brute(oracle, rng):
    ref = rng.uniform(-10, 10, size=(4000, 3)).astype(np.float32)
    ref[100:130] = ref[99]
    qry = rng.uniform(-10, 10, size=(500, 3)).astype(np.float32)
    c = oracle.Cloud(ref)
    for k, cap in ((1, np.inf), (6, np.inf), (6, 0.49), (8, 4.0)):
        i1, d1 = oracle.knn(c, qry, k, cap, kdtree=True)
        i2, d2 = oracle.knn(c, qry, k, cap, kdtree=False)
        assert np.array_equal(i1, i2) and np.array_equal(d1.view(np.uint32), d2.view(np.uint32))


def test_knn_matches_scipy_away_from_ties(oracle, rng):
    ref = rng.uniform(-5, 5, size=(3000, 3)).astype(np.float32)
    qry = rng.uniform(-5, 5, size=(400, 3)).astype(np.float32)
    idx, d2 = oracle.knn(oracle.Cloud(ref), qry, 6, np.inf, kdtree=True)
    dd, ii = cKDTree(ref.astype(np.float64)).query(qry.astype(np.float64), k=6)
    gaps = np.diff(dd, axis=1).min(axis=1) > 1e-4  # compare where the order is unambiguous
    assert gaps.sum() > 300
    assert np.array_equal(idx[gaps], ii[gaps].astype(np.uint32))
    assert np.allclose(np.sqrt(d2[gaps]), dd[gaps], rtol=1e-5, atol=1e-6)


def test_tie_rule_lowest_index(oracle):
    ref = np.array([[1, 0, 0], [0, 1, 0], [-1, 0, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1], [2, 0, 0]], np.float32)
    idx, d2 = oracle.knn(oracle.Cloud(ref), np.zeros((1, 3), np.float32), 4, np.inf)
    assert list(idx[0]) == [0, 1, 2, 3] and (d2[0] == 1.0).all()
    idx, _ = oracle.knn(oracle.Cloud(ref), np.zeros((1, 3), np.float32), 8, np.inf)
    assert list(idx[0]) == [0, 1, 2, 3, 4, 5, 6, 0xFFFFFFFF]


def test_threshold_is_inclusive(oracle):
    cap = np.float32(0.7) * np.float32(0.7)
    x = np.float32(np.sqrt(np.float64(cap)))
    ref = np.array([[x, 0, 0]], np.float32)
    d2 = np.float32(x) * np.float32(x)
    idx, _ = oracle.knn(oracle.Cloud(ref), np.zeros((1, 3), np.float32), 1, cap)
    assert (idx[0, 0] == 0) == bool(d2 <= cap)


def test_eig3_against_numpy(oracle, rng):
    for _ in range(200):
        a = rng.normal(size=(3, 3))
        C = a @ a.T * rng.uniform(1e-4, 10)
        ev, V = oracle.eig3(C)
        w = np.linalg.eigvalsh(C)
        assert np.allclose(ev, w, rtol=1e-12, atol=1e-14 * np.abs(w).max())
        assert np.allclose(C @ V, V * ev, atol=1e-12 * np.abs(w).max())
        assert np.allclose(V.T @ V, np.eye(3), atol=1e-13)
    ev, V = oracle.eig3(np.zeros((3, 3)))
    assert (ev == 0).all() and np.array_equal(V, np.eye(3))


def test_qr_solve_against_numpy(oracle, rng):
    for _ in range(100):
        a = rng.normal(size=(20, 6))
        H = a.T @ a
        b = rng.normal(size=6)
        x, rank = oracle.qr_solve6(H, b)
        assert rank == 6 and np.allclose(x, np.linalg.solve(H, b), rtol=1e-9, atol=1e-11)
    # rank deficient: basic solution solves the consistent system
    a = rng.normal(size=(20, 4)) @ rng.normal(size=(4, 6))
    H = a.T @ a
    b = H @ rng.normal(size=6)
    x, rank = oracle.qr_solve6(H, b)
    assert rank == 4 and np.allclose(H @ x, b, atol=1e-8)


def test_se3_exp_log_roundtrip(oracle, rng):
    from scipy.linalg import expm
    for scale in (1e-9, 1e-5, 1e-2, 1.0, 3.0):
        for _ in range(20):
            eps = rng.normal(size=6) * scale
            R, t = oracle.se3_exp(eps)
            W = np.array([[0, -eps[5], eps[4]], [eps[5], 0, -eps[3]], [-eps[4], eps[3], 0]])
            M = np.zeros((4, 4))
            M[:3, :3], M[:3, 3] = W, eps[:3]
            E = expm(M)
            assert np.allclose(R, E[:3, :3], atol=1e-12) and np.allclose(t, E[:3, 3], atol=1e-12 * max(1, scale))
            if np.linalg.norm(eps[3:]) < 3.0:
                assert np.allclose(oracle.se3_log(R, t), eps, atol=1e-9 * max(1, scale))


def test_pose_ypr_convention(oracle, rng):
    # MRPT: R = Rz(yaw) Ry(pitch) Rx(roll)
    for _ in range(50):
        p = np.r_[rng.normal(size=3), rng.uniform(-3, 3), rng.uniform(-1.5, 1.5), rng.uniform(-3, 3)]
        R, t = oracle.pose_to_Rt(p)
        cy, sy, cp, sp, cr, sr = np.cos(p[3]), np.sin(p[3]), np.cos(p[4]), np.sin(p[4]), np.cos(p[5]), np.sin(p[5])
        Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
        Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
        Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
        assert np.allclose(R, Rz @ Ry @ Rx, atol=1e-14)
        assert np.allclose(oracle.Rt_to_pose(R, t), p, atol=1e-12)


def _plane_pairings(rng, n=300):
    normals = rng.normal(size=(n, 3))
    normals /= np.linalg.norm(normals, axis=1, keepdims=True)
    return rng.uniform(-20, 20, size=(n, 3)), normals


def test_gn_point2plane_recovers_known_pose(oracle, rng):
    P, N = _plane_pairings(rng)
    truth = np.array([0.3, -0.2, 0.05, 0.03, 0.01, -0.02])
    R, t = oracle.pose_to_Rt(truth)
    Cc = P @ R.T + t + np.cross(N, rng.normal(size=N.shape))  # any point of the plane through T p
    Rg, tg, it = oracle.gn_point2plane(P, Cc, N, np.eye(3), np.zeros(3))
    assert it <= 8
    assert np.allclose(oracle.Rt_to_pose(Rg, tg), truth, atol=1e-10)


def test_gn_point2point_and_horn_recover_known_pose(oracle, rng):
    P = rng.uniform(-10, 10, size=(200, 3))
    truth = np.array([1.0, -2.0, 0.5, 0.4, -0.1, 0.2])
    R, t = oracle.pose_to_Rt(truth)
    Q = P @ R.T + t
    Rg, tg, _ = oracle.gn_point2point(P, Q, np.eye(3), np.zeros(3))
    assert np.allclose(oracle.Rt_to_pose(Rg, tg), truth, atol=1e-9)
    Rh, th, used = oracle.horn(P, Q, oracle.default_params())
    assert used == 200 and np.allclose(oracle.Rt_to_pose(Rh, th), truth, atol=1e-9)
    # scale-outlier rule (row M): a pair stretched by 1.5 is dropped, the answer stays exact
    Q2 = Q.copy()
    Q2[0] = Q.mean(0) + 1.5 * (Q[0] - Q.mean(0))
    Rh, th, used = oracle.horn(P, Q2, oracle.default_params())
    assert used == 199
    est = oracle.Rt_to_pose(Rh, th)
    assert np.allclose(est[3:], truth[3:], atol=1e-9)   # rotation: the stretched pair is out
    assert np.allclose(est[:3], truth[:3], atol=5e-2)   # translation uses the centroids of ALL pairs (A.10)


def test_icp_known_transform_c1(oracle):
    """BASELINE config C1 (CPU): noise-free known rigid transform is recovered."""
    from mola_fe_lidar_b200 import scene
    A, B, pose = scene.make_pair_c1(seed=1, n=6000, sigma=0.0)
    r = oracle.icp_align(oracle.Cloud(A), oracle.Cloud(B), np.zeros(6), oracle.default_params(), kdtree=True)
    assert r["termination_reason"] == 4  # Stalled
    assert np.abs(r["pose"][:3] - pose[:3]).max() < 2e-3 and np.abs(r["pose"][3:] - pose[3:]).max() < 2e-4
    assert r["quality"] > 0.9 and not r["cov_singular"]
    assert np.allclose(r["cov"], r["cov"].T, rtol=1e-6) and (np.linalg.eigvalsh(r["cov"]) > 0).all()
    # brute-force search gives the same registration, bit for bit
    r2 = oracle.icp_align(oracle.Cloud(A[:1500]), oracle.Cloud(B[:1500]), np.zeros(6), oracle.default_params(), kdtree=False)
    r3 = oracle.icp_align(oracle.Cloud(A[:1500]), oracle.Cloud(B[:1500]), np.zeros(6), oracle.default_params(), kdtree=True)
    assert np.array_equal(r2["pose"], r3["pose"]) and r2["n_iterations"] == r3["n_iterations"]


def test_icp_edge_cases(oracle, rng):
    prm = oracle.default_params()
    A = rng.uniform(-5, 5, size=(500, 3)).astype(np.float32)
    r = oracle.icp_align(oracle.Cloud(A), oracle.Cloud(A + np.float32(100)), np.zeros(6), prm)
    assert r["termination_reason"] == 1 and r["n_iterations"] == 0 and r["quality"] == 0 and r["cov_singular"]
    e = np.zeros((0, 3), np.float32)
    assert oracle.icp_align(oracle.Cloud(e), oracle.Cloud(A), np.zeros(6), prm)["termination_reason"] == 1
    assert oracle.icp_align(oracle.Cloud(A), oracle.Cloud(e), np.zeros(6), prm)["termination_reason"] == 1
    r = oracle.icp_align(oracle.Cloud(A), oracle.Cloud(A), np.zeros(6), oracle.default_params(max_iterations=0))
    assert r["termination_reason"] == 3 and r["n_iterations"] == 0


def test_covariance_matches_analytic_jacobian(oracle, rng):
    from mola_fe_lidar_b200 import scene
    A, B, pose = scene.make_pair_c1(seed=2, n=3000, sigma=0.0)
    ca, cb = oracle.Cloud(A), oracle.Cloud(B)
    prm = oracle.default_params()
    r = oracle.icp_align(ca, cb, np.zeros(6), prm, kdtree=True)
    # rebuild J analytically (central differences in float128-free numpy) from the final pairings
    m = oracle.match_point2plane(ca, cb, r["R"], r["t"], prm, kdtree=True)
    sel = m["paired"].astype(bool)
    P, Cc, N = B[sel].astype(np.float64), m["centroid"][sel], m["normal"][sel]

    def resid(x):
        R, t = oracle.pose_to_Rt(x)
        return np.einsum("ij,ij->i", N, P @ R.T + t - Cc)
    x0 = oracle.Rt_to_pose(r["R"], r["t"])
    J = np.stack([(resid(x0 + 1e-6 * np.eye(6)[j]) - resid(x0 - 1e-6 * np.eye(6)[j])) / 2e-6 for j in range(6)], 1)
    # the oracle's final pairings were made one iteration earlier; the set is the same at convergence
    if m["n"] == r["n_pairings"]:
        assert np.allclose(np.linalg.inv(J.T @ J), r["cov"], rtol=2e-3, atol=1e-12)


def test_voxel_decimation_properties(oracle, rng):
    pts = rng.uniform(-10, 10, size=(20000, 3)).astype(np.float32)
    keep, xyz = oracle.voxel_decimate(pts, 1.0)
    keys = np.floor(pts / np.float32(1.0)).astype(np.int64)
    _, first = np.unique(keys, axis=0, return_index=True)
    assert np.array_equal(keep, np.sort(first).astype(np.uint32))
    assert np.array_equal(xyz, pts[keep])
    keep2, xyz2 = oracle.voxel_decimate(pts, 1.0, use_average=True)
    assert np.array_equal(keep2, keep)
    k0 = keys[keep[0]]
    mean = pts[(keys == k0).all(1)].astype(np.float64).mean(0)
    assert np.allclose(xyz2[0], mean, atol=1e-5)
    assert len(oracle.voxel_decimate(np.zeros((0, 3), np.float32), 1.0)[0]) == 0


def test_golden_vectors(oracle):
    """Fixtures written by tests/golden/make_golden.py from the independent
    implementations (scipy / numpy); the oracle must keep reproducing them."""
    g = np.load(os.path.join(GOLDEN, "knn_c1.npz"))
    idx, d2 = oracle.knn(oracle.Cloud(g["ref"]), g["qry"], 6, np.float32(0.49), kdtree=True)
    assert np.array_equal(idx, g["idx"]) and np.array_equal(d2.view(np.uint32), g["d2_bits"])
    meta = json.load(open(os.path.join(GOLDEN, "icp_c1.json")))
    from mola_fe_lidar_b200 import scene
    A, B, _ = scene.make_pair_c1(seed=meta["seed"], n=meta["n"], sigma=meta["sigma"])
    r = oracle.icp_align(oracle.Cloud(A), oracle.Cloud(B), np.zeros(6), oracle.default_params(), kdtree=True)
    assert np.allclose(r["pose"], meta["pose"], atol=1e-9)
    assert r["n_iterations"] == meta["n_iterations"] and r["n_pairings"] == meta["n_pairings"]
    assert abs(r["quality"] - meta["quality"]) < 1e-12


def test_filter_edges_planes_known_answers(oracle, rng):
    """A.13 on shapes whose class is known: a vertical wall is 'planes', a horizontal one is dropped (ground-like),
    an isotropic blob is 'edges', a thin line is neither (e2 >> e0 but e1 ~ e0), a voxel below the point minimum is
    only in full_decim; decimation takes every d-th point of a voxel in ascending original index."""
    def one(pts, **kw):
        f, nv = oracle.filter_edges_planes(np.asarray(pts, dtype=np.float32), **kw)
        return f, nv
    wall = np.c_[rng.uniform(0.01, 0.99, 200), 0.5 + rng.normal(0, 1e-3, 200), rng.uniform(0.01, 0.99, 200)]
    f, nv = one(wall, voxel_filter_decimation=1, full_pointcloud_decimation=1)
    assert nv == 1 and (f == 6).all()
    ground = wall[:, [0, 2, 1]]
    f, nv = one(ground, voxel_filter_decimation=1, full_pointcloud_decimation=1)
    assert nv == 0 and (f == 4).all()
    blob = rng.uniform(0.01, 0.99, (200, 3))
    f, nv = one(blob, voxel_filter_decimation=1, full_pointcloud_decimation=1)
    assert nv == 1 and (f == 5).all()
    line = np.c_[rng.uniform(0.01, 0.99, 200), 0.5 + rng.normal(0, 1e-3, 200), 0.5 + rng.normal(0, 1e-3, 200)]
    f, nv = one(line, voxel_filter_decimation=1, full_pointcloud_decimation=1)
    assert nv == 0 and (f == 4).all()
    f, nv = one(blob[:4], voxel_filter_decimation=1, full_pointcloud_decimation=1)
    assert nv == 0 and (f == 4).all()
    f, nv = one(blob, voxel_filter_decimation=10, full_pointcloud_decimation=7)
    assert np.array_equal(np.flatnonzero(f & 1), np.arange(0, 200, 10))
    assert np.array_equal(np.flatnonzero(f & 4), np.arange(0, 200, 7))
    f, nv = one(np.array([[np.nan, 0, 0], [0.5, 0.5, 0.5]]))
    assert f[0] == 0 and f[1] == 4


def test_filter_edges_planes_matches_independent_golden(oracle):
    """tests/golden/edges_planes_c1.npz: the voxel classes from numpy's eigh on np.unique voxels, kept where every gate
    is at least 2 % away from its threshold (make_golden.py) -- the oracle's hash/sort + Jacobi path must agree."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "edges_planes_c1.npz"))
    f, _ = oracle.filter_edges_planes(g["pts"], voxel_filter_decimation=1, full_pointcloud_decimation=1)
    cls_of_point = np.where(f & 1, 1, np.where(f & 2, 2, 0))
    want = g["voxel_class"][g["voxel_of_point"]]
    sure = g["voxel_sure"][g["voxel_of_point"]]
    assert sure.sum() > 3000 and (want[sure] == 1).any() and (want[sure] == 2).any()
    assert np.array_equal(cls_of_point[sure], want[sure])
    assert (f & 4).all()
