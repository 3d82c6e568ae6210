"""The selectable solver / matcher classes of the ICP YAML (`solvers: - class:`,
`matchers: - class:`, reference seam src/LidarOdometry.cpp:80-84) on the device
vs the oracle: Solver_Horn (row N) with the pairings-weight rules (row M), and
Matcher_Points_DistanceThreshold as the ICP matcher, in all four combinations."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL_T, TOL_R = 1e-5, 1e-6
GN, HORN = 0, 1
P2PL, P2P = 0, 1


def _both(capi, oracle, **kw):
    return capi.default_params(**kw), oracle.default_params(**kw)


def _align(capi, oracle, A, B, guess, **kw):
    gp, op = _both(capi, oracle, **kw)
    icp = capi.ICP(gp, device=0)
    ga, gb = icp.upload(A), icp.upload(B)
    g = icp.align(ga, gb, guess)
    o = oracle.icp_align(oracle.Cloud(A), oracle.Cloud(B), guess, op, kdtree=True)
    ga.free(), gb.free()
    icp.close()
    return g, o


def _same(g, o, tol_t=TOL_T, tol_r=TOL_R):
    assert g["termination_reason"] == o["termination_reason"], (g, o)
    assert g["n_iterations"] == o["n_iterations"]
    assert g["n_pairings"] == o["n_pairings"]
    assert g["quality"] == o["quality"]
    assert np.abs(g["pose"][:3] - o["pose"][:3]).max() < tol_t, (g["pose"], o["pose"])
    assert np.abs(g["pose"][3:] - o["pose"][3:]).max() < tol_r, (g["pose"], o["pose"])
    if not o["cov_singular"]:
        scale = np.sqrt(np.outer(np.diag(o["cov"]), np.diag(o["cov"])))
        assert (np.abs(g["cov"] - o["cov"]) / scale).max() < 1e-4


def test_point2point_matcher_bit_exact(capi, oracle):
    """1-NN with the strict d2 < thr^2 gate: indices and pairing flags bit-exact."""
    from mola_fe_lidar_b200 import scene
    A, B, pose = scene.make_pair_c1(seed=2, n=20000, sigma=0.01)
    gp, op = _both(capi, oracle, matcher_kind=P2P, distance_threshold=0.25)
    icp = capi.ICP(gp, device=0)
    ga, gb = icp.upload(A), icp.upload(B)
    for guess in (np.zeros(6), pose):
        m = icp.match(ga, gb, guess)
        R, t = oracle.pose_to_Rt(guess)
        n, nn, d2 = oracle.match_points(oracle.Cloud(A), oracle.Cloud(B), R, t, 0.25, kdtree=True)
        assert m["n"] == n > 1000
        assert np.array_equal(m["nn_idx"][:, 0], nn)
        assert np.array_equal(m["paired"].astype(bool), nn != 0xFFFFFFFF)
        sel = nn != 0xFFFFFFFF
        assert np.array_equal(m["centroid"][sel], A[nn[sel]].astype(np.float64))
    ga.free(), gb.free()
    icp.close()


@pytest.mark.parametrize("solver,matcher", [(GN, P2P), (HORN, P2P), (HORN, P2PL)])
def test_c1_pair_all_combinations(capi, oracle, solver, matcher):
    from mola_fe_lidar_b200 import scene
    A, B, pose = scene.make_pair_c1(seed=1, n=20000, sigma=0.01)
    g, o = _align(capi, oracle, A, B, np.zeros(6), solver_kind=solver, matcher_kind=matcher)
    _same(g, o)
    assert g["n_iterations"] >= 2
    # every combination still lands near the true transform
    assert np.abs(g["pose"][:3] - pose[:3]).max() < 0.05 and np.abs(g["pose"][3:] - pose[3:]).max() < 0.01


def test_horn_weight_rules(capi, oracle):
    """row M: scale-outlier detector on/off and the robust kernel change the
    result identically on both sides."""
    from mola_fe_lidar_b200 import scene
    A, B, pose = scene.make_pair_c1(seed=3, n=12000, sigma=0.02)
    g0 = pose * 0.5
    res = []
    for kw in (dict(use_scale_outlier_detector=0), dict(use_scale_outlier_detector=1, scale_outlier_threshold=1.05),
               dict(use_robust_kernel=1, robust_kernel_param=np.deg2rad(0.1), robust_kernel_scale=400.0)):
        g, o = _align(capi, oracle, A, B, g0, solver_kind=HORN, matcher_kind=P2P, **kw)
        # the robust weight goes through acos(): device and libm differ by ulps
        _same(g, o, tol_t=2e-5 if "use_robust_kernel" in kw else TOL_T, tol_r=2e-6 if "use_robust_kernel" in kw else TOL_R)
        res.append(g["pose"])
    assert not np.allclose(res[0], res[1], atol=1e-9) or not np.allclose(res[0], res[2], atol=1e-9)


def test_horn_solver_error_and_no_pairings(capi, oracle, rng):
    # two pairings only: fewer than 3 usable pairs -> SolverError, pose untouched
    A = np.array([[0, 0, 0], [1, 0, 0], [50, 50, 50]], dtype=np.float32)
    B = np.array([[0.01, 0, 0], [1.01, 0, 0]], dtype=np.float32)
    g, o = _align(capi, oracle, A, B, np.zeros(6), solver_kind=HORN, matcher_kind=P2P)
    assert g["termination_reason"] == o["termination_reason"] == 2  # SolverError
    assert g["n_iterations"] == o["n_iterations"] == 0 and np.all(g["pose"] == 0)
    far = (rng.uniform(-1, 1, (100, 3)) + 500).astype(np.float32)
    g, o = _align(capi, oracle, A, far, np.zeros(6), solver_kind=HORN, matcher_kind=P2P)
    assert g["termination_reason"] == o["termination_reason"] == 1  # NoPairings


def test_horn_from_yaml_class_names(capi, oracle):
    """`class: mp2p_icp::Solver_Horn` / `Matcher_Points_DistanceThreshold` select the device paths."""
    import os
    from mola_fe_lidar_b200 import lidar_odometry, scene
    txt = open(os.path.join(lidar_odometry.PARAMS_DIR, "icp-settings-regular.yaml")).read()
    txt = txt.replace("mp2p_icp::Solver_GaussNewton", "mp2p_icp::Solver_Horn")
    txt = txt.replace("mp2p_icp::Matcher_Point2Plane", "mp2p_icp::Matcher_Points_DistanceThreshold")
    p = capi.params_from_yaml(txt)
    assert p.solver_kind == HORN and p.matcher_kind == P2P
    icp = capi.ICP(yaml_text=txt, device=0)
    A, B, pose = scene.make_pair_c1(seed=5, n=8000, sigma=0.0)
    ga, gb = icp.upload(A), icp.upload(B)
    g = icp.align(ga, gb, np.zeros(6))
    op = oracle.default_params(solver_kind=HORN, matcher_kind=P2P)
    for k in ("distance_threshold", "max_iterations", "use_scale_outlier_detector", "scale_outlier_threshold"):
        setattr(op, k, getattr(p, k))
    o = oracle.icp_align(oracle.Cloud(A), oracle.Cloud(B), np.zeros(6), op, kdtree=True)
    _same(g, o)
    ga.free(), gb.free()
    icp.close()


def test_batch_with_horn(capi, oracle, rng):
    from mola_fe_lidar_b200 import scene
    A, B, pose = scene.make_pair_c1(seed=7, n=6000, sigma=0.005)
    gp, op = _both(capi, oracle, solver_kind=HORN, matcher_kind=P2P)
    icp = capi.ICP(gp, device=0)
    ga, gb = icp.upload(A), icp.upload(B)
    guesses = pose + np.c_[rng.normal(0, 0.03, (4, 3)), rng.normal(0, 0.005, (4, 1)), np.zeros((4, 2))]
    out = icp.align_batch([ga] * 4, [gb] * 4, guesses)
    for i in range(4):
        o = oracle.icp_align(oracle.Cloud(A), oracle.Cloud(B), guesses[i], op, kdtree=True)
        _same(out[i], o)
    ga.free(), gb.free()
    icp.close()
