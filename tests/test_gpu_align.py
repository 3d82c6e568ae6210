"""Whole registrations on the device vs the oracle (rows A, G-P): final pose
within 1e-5 m / 1e-6 rad (BASELINE.json north_star), same iteration count,
termination reason, pairing count and quality; covariance to 1e-6 relative."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL_T, TOL_R = 1e-5, 1e-6


def _assert_same(g, o, cov=True):
    assert np.abs(g["pose"][:3] - o["pose"][:3]).max() < TOL_T, (g["pose"], o["pose"])
    assert np.abs(g["pose"][3:] - o["pose"][3:]).max() < TOL_R, (g["pose"], o["pose"])
    assert g["termination_reason"] == o["termination_reason"]
    assert g["n_iterations"] == o["n_iterations"]
    assert g["n_pairings"] == o["n_pairings"]
    assert g["quality"] == o["quality"]
    if cov and not o["cov_singular"]:
        assert not g["cov_singular"]
        scale = np.sqrt(np.outer(np.diag(o["cov"]), np.diag(o["cov"])))
        assert np.abs(g["cov"] - o["cov"]).max() / scale.max() < 1e-5
        assert (np.abs(g["cov"] - o["cov"]) / scale).max() < 1e-4


def _run(icp, oracle, A, B, guess, params=None):
    g_a, g_b = icp.upload(A), icp.upload(B)
    g = icp.align(g_a, g_b, guess)
    o = oracle.icp_align(oracle.Cloud(A), oracle.Cloud(B), guess,
                         oracle.default_params() if params is None else params, kdtree=True)
    g_a.free(), g_b.free()
    return g, o


@pytest.mark.parametrize("sigma", [0.0, 0.01])
def test_c1_known_transform(icp, oracle, sigma):
    """BASELINE config C1: two 20k-pt clouds, known rigid transform."""
    from mola_fe_lidar_b200 import scene
    A, B, pose = scene.make_pair_c1(seed=1, n=20000, sigma=sigma)
    g, o = _run(icp, oracle, A, B, np.zeros(6))
    _assert_same(g, o)
    assert g["quality"] > 0.5


def test_identity(icp, oracle):
    from mola_fe_lidar_b200 import scene
    A, _, _ = scene.make_pair_c1(seed=2, n=8000)
    g, o = _run(icp, oracle, A, A, np.zeros(6))
    _assert_same(g, o)
    assert np.abs(g["pose"]).max() < 2e-3 and g["quality"] == 1.0
    assert g["termination_reason"] == 4


def test_no_pairings(icp, oracle, rng):
    A = rng.uniform(-5, 5, size=(2000, 3)).astype(np.float32)
    B = A + np.float32(100.0)
    g, o = _run(icp, oracle, A, B, np.zeros(6))
    assert g["termination_reason"] == 1 == o["termination_reason"]
    assert g["n_iterations"] == 0 and g["quality"] == 0.0 and g["cov_singular"] == 1
    assert np.array_equal(g["pose"], o["pose"])


def test_scan_to_scan_120k(icp, oracle):
    """One step of BASELINE config C2 (raw 120k-pt scans)."""
    from mola_fe_lidar_b200 import scene
    scans, poses = scene.make_sequence(2, seed=1)
    g, o = _run(icp, oracle, scans[0], scans[1], np.zeros(6))
    _assert_same(g, o)


def test_batch_equals_single(icp, oracle, rng):
    """align_batch (Monte-Carlo guesses, LidarOdometry.cpp:775-787) == independent aligns."""
    from mola_fe_lidar_b200 import scene
    A, B, pose = scene.make_pair_c1(seed=3, n=6000, sigma=0.005)
    C2, D2, _ = scene.make_pair_c1(seed=4, n=3000, sigma=0.0)
    g_a, g_b, g_c, g_d = icp.upload(A), icp.upload(B), icp.upload(C2), icp.upload(D2)
    guesses = pose + np.c_[rng.normal(0, 0.05, (5, 3)), rng.normal(0, 0.01, (5, 1)), np.zeros((5, 2))]
    guesses = np.concatenate([guesses, np.zeros((1, 6))])
    froms = [g_a] * 5 + [g_c]
    tos = [g_b] * 5 + [g_d]
    batch = icp.align_batch(froms, tos, guesses)
    for i in range(6):
        single = icp.align(froms[i], tos[i], guesses[i])
        # the number of CTAs per job (hence the fixed summation order of the
        # moment partials) depends on how many jobs share a launch: same
        # correspondences, sums equal to rounding
        assert np.abs(batch[i]["pose"] - single["pose"]).max() < 1e-9
        assert np.allclose(batch[i]["cov"], single["cov"], rtol=1e-6, atol=1e-12)
        for key in ("quality", "n_iterations", "termination_reason", "n_pairings"):
            assert batch[i][key] == single[key], key
    o = oracle.icp_align(oracle.Cloud(A), oracle.Cloud(B), guesses[0], oracle.default_params())
    _assert_same(batch[0], o)


def test_run_to_run_bit_reproducible(icp):
    from mola_fe_lidar_b200 import scene
    A, B, _ = scene.make_pair_c1(seed=5, n=10000, sigma=0.01)
    g_a, g_b = icp.upload(A), icp.upload(B)
    r1 = icp.align(g_a, g_b, np.zeros(6))
    g_a2, g_b2 = icp.upload(A), icp.upload(B)
    r2 = icp.align(g_a2, g_b2, np.zeros(6))
    assert np.array_equal(r1["pose"], r2["pose"]) and np.array_equal(r1["cov"], r2["cov"])


def test_per_call_parameters_keep_the_objects_matchers(icp, oracle):
    """icp->align(from, to, guess, in.icp_params, result) (LidarOdometry.cpp:869-871): the call's
    mp2p_icp::Parameters (iteration budget, step tolerances) on the OBJECT's matchers / solvers / quality
    evaluators -- b200icp_align_with equals the oracle run with those Parameters, and no second device
    object is involved."""
    from mola_fe_lidar_b200 import scene
    A, B, _ = scene.make_pair_c1(seed=6, n=8000, sigma=0.01)
    g_a, g_b = icp.upload(A), icp.upload(B)
    full = icp.align(g_a, g_b, np.zeros(6))
    assert full["n_iterations"] > 2
    for kw in ({"max_iterations": 2}, {"min_abs_step_trans": 1e-2, "min_abs_step_rot": 1e-2}):
        g = icp.align_with(g_a, g_b, np.zeros(6), **kw)
        prm = oracle.default_params()
        for k, v in kw.items():
            setattr(prm, k, v)
        o = oracle.icp_align(oracle.Cloud(A), oracle.Cloud(B), np.zeros(6), prm, kdtree=True)
        _assert_same(g, o)
        assert g["n_iterations"] < full["n_iterations"]
    # the object's own parameters are untouched
    again = icp.align(g_a, g_b, np.zeros(6))
    assert np.array_equal(again["pose"], full["pose"]) and again["n_iterations"] == full["n_iterations"]


def test_varying_cloud_sizes_back_to_back(icp, oracle):
    """Real scans (and decimated ones) change their point count every frame: consecutive registrations of
    clouds of different sizes on one ICP object stay correct whatever launch configuration is reused."""
    from mola_fe_lidar_b200 import scene
    for seed, n in ((11, 5000), (12, 5100), (13, 4700), (14, 9000), (15, 5000), (16, 300)):
        A, B, _ = scene.make_pair_c1(seed=seed, n=n, sigma=0.005)
        g, o = _run(icp, oracle, A, B, np.zeros(6))
        _assert_same(g, o)


def test_period_two_fast_forward_matches_the_full_loop(capi, oracle):
    """Registrations of coarse clouds that end in a cycle between two or three pairing sets run to maxIterations in
    the reference (in this sequence: one 3-cycle, pair 3 -> 4, and one 2-cycle, pair 11 -> 12).  The device recognises
    the cycle (pose and pairing count repeating those of p iterations before, p times in a row, p = 2 .. 6) and jumps
    to the end state: same iteration count, termination reason, pairing count and quality as running the loop out
    (B200ICP_CYCLE=0 in a child process), poses within 1e-10 -- and equal to the oracle's within the stated
    tolerance."""
    import json
    import os
    import subprocess
    import sys
    from mola_fe_lidar_b200 import scene
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    child = r"""
import sys, json
import numpy as np
sys.path.insert(0, %r)
from mola_fe_lidar_b200 import capi, scene
icp = capi.ICP(capi.default_params())
scans, _ = scene.make_sequence(13, seed=1)
dec = [icp.voxel_decimate(icp.upload_raw(s), 1.0) for s in scans]
out, guess = [], np.zeros(6)
for i in range(1, 13):
    r = icp.align(dec[i - 1], dec[i], guess)
    out.append(dict(pose=[float(v).hex() for v in r["pose"]], cov=[float(v).hex() for v in r["cov"].ravel()],
                    q=float(r["quality"]).hex(), it=int(r["n_iterations"]), term=int(r["termination_reason"]),
                    npair=int(r["n_pairings"])))
    guess = np.array([r["pose"][0], r["pose"][1], r["pose"][2], r["pose"][3], 0, 0])
print(json.dumps(out))
""" % root
    runs = {}
    for flag in ("2", "1", "0"):
        env = dict(os.environ, B200ICP_CYCLE=flag)
        p = subprocess.run([sys.executable, "-c", child], capture_output=True, text=True, env=env, timeout=600)
        assert p.returncode == 0, p.stderr[-2000:]
        runs[flag] = json.loads(p.stdout.strip().splitlines()[-1])
    for a, b in zip(runs["2"], runs["0"]):
        assert (a["it"], a["term"], a["npair"], a["q"]) == (b["it"], b["term"], b["npair"], b["q"])
        pa = np.array([float.fromhex(v) for v in a["pose"]]), np.array([float.fromhex(v) for v in b["pose"]])
        assert np.abs(pa[0] - pa[1]).max() < 1e-10
        ca = np.array([float.fromhex(v) for v in a["cov"]]), np.array([float.fromhex(v) for v in b["cov"]])
        assert np.allclose(ca[0], ca[1], rtol=1e-6, atol=1e-14)
    assert runs["1"] == runs["0"]  # bitwise-only recognition never changes a bit
    runs["1"] = runs["2"]
    # the sequence must actually contain registrations that exhaust the iteration budget
    assert sum(r["term"] == 3 and r["it"] == 100 for r in runs["1"]) >= 2
    # and the oracle agrees on one of them
    scans, _ = scene.make_sequence(13, seed=1)
    k = next(i for i, r in enumerate(runs["1"]) if r["term"] == 3)
    dec = [oracle.voxel_decimate(s, 1.0)[1] for s in scans[k:k + 2]]
    prev = runs["1"][k - 1]["pose"] if k else None
    guess = np.zeros(6) if k == 0 else np.array([float.fromhex(v) for v in prev[:4]] + [0.0, 0.0])
    o = oracle.icp_align(oracle.Cloud(dec[0]), oracle.Cloud(dec[1]), guess, oracle.default_params(), kdtree=True)
    g = runs["1"][k]
    assert o["n_iterations"] == g["it"] and o["termination_reason"] == g["term"] and o["n_pairings"] == g["npair"]
    pose = np.array([float.fromhex(v) for v in g["pose"]])
    assert np.abs(pose[:3] - o["pose"][:3]).max() < TOL_T and np.abs(pose[3:] - o["pose"][3:]).max() < TOL_R


def test_concurrent_align_on_one_object(icp, oracle):
    """One shared ICP object, align() called concurrently from the pool threads (LidarOdometry.h:167-172,
    cpp:94-96, 711-729): eight threads register different pairs at once, several rounds; every result must be
    the bits the same call gives alone, and agree with the oracle."""
    import threading
    from mola_fe_lidar_b200 import scene
    pairs = []
    for t in range(8):
        A, B, _ = scene.make_pair_c1(seed=20 + t, n=4000 + 700 * t, sigma=0.005)
        pairs.append((A, B, icp.upload(A), icp.upload(B)))
    alone = [icp.align(p[2], p[3], np.zeros(6)) for p in pairs]
    results, errors = [[None] * 4 for _ in range(8)], []

    def work(t):
        try:
            for rep in range(4):
                results[t][rep] = icp.align(pairs[t][2], pairs[t][3], np.zeros(6))
        except Exception as e:  # pragma: no cover
            errors.append(e)

    th = [threading.Thread(target=work, args=(t,)) for t in range(8)]
    [x.start() for x in th]
    [x.join() for x in th]
    assert not errors, errors
    for t in range(8):
        for rep in range(4):
            r = results[t][rep]
            assert np.array_equal(r["pose"], alone[t]["pose"]) and np.array_equal(r["cov"], alone[t]["cov"])
            for key in ("quality", "n_iterations", "termination_reason", "n_pairings"):
                assert r[key] == alone[t][key]
    o = oracle.icp_align(oracle.Cloud(pairs[3][0]), oracle.Cloud(pairs[3][1]), np.zeros(6), oracle.default_params(),
                         kdtree=True)
    _assert_same(alone[3], o)
