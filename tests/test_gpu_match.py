"""Matcher_Point2Plane on the device vs the oracle at a fixed pose: pairing
decisions, neighbour lists, plane centroids and normals bit-exact
(rows H, J)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _compare(icp, oracle, glob, loc, pose, params=None):
    op = oracle.default_params() if params is None else params
    g_g, g_l = icp.upload(glob), icp.upload(loc)
    m = icp.match(g_g, g_l, pose)
    R, t = oracle.pose_to_Rt(pose)
    o = oracle.match_point2plane(oracle.Cloud(glob), oracle.Cloud(loc), R, t, op, kdtree=True)
    assert np.array_equal(m["nn_cnt"], o["nn_cnt"])
    assert np.array_equal(m["nn_idx"], o["nn_idx"])
    assert np.array_equal(m["paired"], o["paired"]), \
        f"pairing decisions differ on {(m['paired'] != o['paired']).sum()} points"
    assert m["n"] == o["n"]
    sel = o["paired"].astype(bool)
    assert np.array_equal(m["centroid"][sel].view(np.uint64), o["centroid"][sel].view(np.uint64))
    assert np.array_equal(m["normal"][sel].view(np.uint64), o["normal"][sel].view(np.uint64))
    g_g.free(), g_l.free()
    return m


def test_match_c1_pair(icp, oracle):
    from mola_fe_lidar_b200 import scene
    A, B, pose = scene.make_pair_c1(seed=1, n=20000, sigma=0.01)
    m = _compare(icp, oracle, A, B, np.zeros(6))
    assert m["n"] > 1000
    _compare(icp, oracle, A, B, pose)


def test_match_structured_planes(icp, oracle, rng):
    # three orthogonal noisy planes: exercises the eigen gate on both sides
    n = 4000
    a = np.c_[rng.uniform(-5, 5, n), rng.uniform(-5, 5, n), rng.normal(0, 0.01, n)]
    b = np.c_[rng.uniform(-5, 5, n), rng.normal(0, 0.01, n) + 5, rng.uniform(0, 5, n)]
    c = np.c_[rng.normal(0, 0.3, n) - 5, rng.uniform(-5, 5, n), rng.uniform(0, 5, n)]  # thick: rejected
    glob = np.concatenate([a, b, c]).astype(np.float32)
    loc = (glob[::3] + rng.normal(0, 0.02, size=glob[::3].shape)).astype(np.float32)
    m = _compare(icp, oracle, glob, loc, np.array([0.05, -0.03, 0.02, 0.01, 0.0, 0.005]))
    assert 0 < m["n"] < len(loc)


def test_match_too_few_neighbours_and_no_pairings(icp, oracle, rng):
    glob = rng.uniform(-50, 50, size=(300, 3)).astype(np.float32)  # sparse: < 3 neighbours
    loc = rng.uniform(-50, 50, size=(500, 3)).astype(np.float32)
    m = _compare(icp, oracle, glob, loc, np.zeros(6))
    assert m["n"] == 0


def test_match_full_scan(icp, oracle):
    from mola_fe_lidar_b200 import scene
    scans, poses = scene.make_sequence(2, seed=3)
    _compare(icp, oracle, scans[0], scans[1], scene.relative_pose6(poses[0], poses[1]))
