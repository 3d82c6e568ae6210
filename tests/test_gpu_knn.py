"""GPU kNN (grid index) vs the oracle's exact search: indices and float d2
bit-exact under the (d2, index) tie rule of SURVEY Appendix A.3/A.4."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _check(icp, oracle, ref, qry, k, max_dist, pose=None, kdtree=False):
    ref = np.ascontiguousarray(ref, dtype=np.float32)
    qry = np.ascontiguousarray(qry, dtype=np.float32)
    g_ref, g_q = icp.upload(ref), icp.upload(qry)
    idx, d2 = icp.knn(g_ref, g_q, k, max_dist, pose6=pose)
    o_ref = oracle.Cloud(ref)
    if pose is not None:
        R, t = oracle.pose_to_Rt(pose)
        q_t = oracle.transform_points(R, t, qry)
    else:
        q_t = qry
    cap = np.float32(max_dist) * np.float32(max_dist)
    oidx, od2 = oracle.knn(o_ref, q_t, k, cap, kdtree=kdtree)
    assert np.array_equal(idx, oidx), f"{(idx != oidx).any(axis=1).sum()} of {len(idx)} queries differ"
    assert np.array_equal(d2.view(np.uint32), od2.view(np.uint32))
    g_ref.free(), g_q.free()
    return idx


@pytest.mark.parametrize("k", [1, 3, 6, 8])
def test_uniform_random(icp, oracle, rng, k):
    ref = rng.uniform(-5, 5, size=(6000, 3))
    qry = rng.uniform(-5.5, 5.5, size=(3000, 3))
    _check(icp, oracle, ref, qry, k, 0.7)


def test_small_radius_and_large_radius(icp, oracle, rng):
    ref = rng.uniform(-3, 3, size=(4000, 3))
    qry = rng.uniform(-3, 3, size=(1500, 3))
    _check(icp, oracle, ref, qry, 6, 0.1)   # quality-evaluator radius
    _check(icp, oracle, ref, qry, 6, 1.5)   # more than one ring of cells


def test_clustered_and_duplicates(icp, oracle, rng):
    centres = rng.uniform(-20, 20, size=(30, 3))
    ref = (centres[rng.integers(0, 30, 5000)] + rng.normal(0, 0.2, size=(5000, 3))).astype(np.float32)
    ref[100:400] = ref[50]                      # 300 exact duplicates of one point
    ref[1000:1100] = ref[900:1000]              # 100 duplicated pairs
    qry = np.concatenate([ref[::7], centres.astype(np.float32)])
    idx = _check(icp, oracle, ref, qry, 6, 0.7)
    # duplicates resolve to the LOWEST indices
    q50 = np.where((qry == ref[50]).all(axis=1))[0][0]
    assert list(idx[q50]) == [50, 100, 101, 102, 103, 104]


def test_exact_tie_lattice(icp, oracle):
    g = np.arange(-6, 7, dtype=np.float32) * 0.5
    ref = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    # cell centres: 8 equidistant corners each; lattice points themselves: 6 ties at 0.5
    qry = np.concatenate([ref[::5] + np.float32(0.25), ref[::11]])
    _check(icp, oracle, ref, qry, 6, 0.7)
    _check(icp, oracle, ref, qry, 1, 0.7)


def test_threshold_edge_kept(icp, oracle):
    # neighbour at exactly d2 == cap is kept ('>' cut, A.5)
    cap = np.float32(0.7) * np.float32(0.7)
    d = np.sqrt(np.float64(cap)).astype(np.float32)
    ref = np.array([[d, 0, 0], [0, 0, 2.0], [5, 5, 5]], dtype=np.float32)
    qry = np.zeros((1, 3), dtype=np.float32)
    idx = _check(icp, oracle, ref, qry, 6, 0.7)
    d2 = np.float32(d) * np.float32(d)
    assert (idx[0, 0] == 0) == (d2 <= cap)


def test_pose_and_nonfinite(icp, oracle, rng):
    ref = rng.uniform(-8, 8, size=(5000, 3)).astype(np.float32)
    qry = rng.uniform(-8, 8, size=(2000, 3)).astype(np.float32)
    ref[17] = np.nan
    ref[99, 1] = np.inf
    qry[5] = np.nan
    pose = np.array([0.3, -0.2, 0.05, 0.4, 0.02, -0.01])
    idx = _check(icp, oracle, ref, qry, 6, 0.7, pose=pose)
    assert (idx[5] == 0xFFFFFFFF).all()
    assert not np.isin(idx, [17, 99]).any()


def test_empty_and_tiny(icp, oracle):
    ref = np.zeros((0, 3), dtype=np.float32)
    qry = np.ones((4, 3), dtype=np.float32)
    g_ref, g_q = icp.upload(ref), icp.upload(qry)
    idx, d2 = icp.knn(g_ref, g_q, 6, 0.7)
    assert (idx == 0xFFFFFFFF).all() and np.isinf(d2).all()
    idx, d2 = icp.knn(g_q, g_ref, 6, 0.7)
    assert idx.shape == (0, 6)
    _check(icp, oracle, np.ones((2, 3)), qry, 6, 0.7)   # fewer than k points


def test_far_coordinates(icp, oracle, rng):
    # UTM-like offsets: cells are keyed relative to the bbox, not the origin
    off = np.array([4.3e4, -1.2e4, 310.0])
    ref = (rng.uniform(-30, 30, size=(6000, 3)) + off).astype(np.float32)
    qry = (rng.uniform(-30, 30, size=(2000, 3)) + off).astype(np.float32)
    _check(icp, oracle, ref, qry, 6, 0.7)


def test_lidar_scan_full_size(icp, oracle):
    """BASELINE kNN config (120k, 120k, k=6 and k=1) against the oracle kd-tree."""
    from mola_fe_lidar_b200 import scene
    scans, poses = scene.make_sequence(2, seed=2)
    guess = scene.relative_pose6(poses[0], poses[1])
    _check(icp, oracle, scans[0], scans[1], 6, 0.7, pose=guess, kdtree=True)
    _check(icp, oracle, scans[0], scans[1], 1, 0.1, pose=guess, kdtree=True)


@pytest.mark.parametrize("k", [1, 6, 8])
def test_uncapped(icp, oracle, rng, k):
    """max_dist = +inf (SURVEY 8d 'radius cap 0.7 m and uncapped'): rows the radius-capped pass leaves incomplete are
    completed over the whole reference cloud; queries far outside the cloud, duplicates and exact ties included."""
    centres = rng.uniform(-20, 20, size=(20, 3))
    ref = (centres[rng.integers(0, 20, 5000)] + rng.normal(0, 0.3, size=(5000, 3))).astype(np.float32)
    ref[200:260] = ref[7]
    qry = np.concatenate([ref[::9], rng.uniform(-60, 60, size=(1500, 3)).astype(np.float32),
                          np.array([[1e4, -1e4, 3e3], [0, 0, 0]], dtype=np.float32)])
    _check(icp, oracle, ref, qry, k, np.inf)
    g = np.arange(-4, 5, dtype=np.float32) * 2.0   # lattice coarser than the index radius: every row needs pass 2
    lat = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    _check(icp, oracle, lat, np.concatenate([lat[::3] + np.float32(1.0), lat[::7]]), k, np.inf)


def test_uncapped_fewer_points_than_k(icp, oracle, rng):
    ref = rng.uniform(-1, 1, size=(4, 3))
    qry = rng.uniform(-30, 30, size=(100, 3))
    _check(icp, oracle, ref, qry, 6, np.inf)
