"""Host-module logic on a box WITHOUT a GPU: the C++ LidarOdometry mirror
(reference src/LidarOdometry.cpp:162-514, 746-849) is driven with a TEST
DOUBLE of the device library (tests/stub/fake_b200icp.c, LD_PRELOADed into a
child process) whose "align" returns a scripted motion.  What is checked is
the front-end control flow around the ICP seam, which must equal the
reference's: time gate, sensor-label filter, constant-velocity guess, twist
update, key-frame rule, factor emission, back-pressure drop rule.  The
numerical path itself is only tested on the GPU (tests/test_gpu_*.py)."""
import json
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STUB_SRC = os.path.join(ROOT, "tests", "stub", "fake_b200icp.c")


@pytest.fixture(scope="module")
def fake_lib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("stub") / "libfake_b200icp.so")
    subprocess.check_call(["gcc", "-shared", "-fPIC", "-O1", "-o", out, STUB_SRC])
    return out


def run_child(fake_lib, body):
    code = textwrap.dedent("""
        import json, sys, ctypes
        import numpy as np
        sys.path.insert(0, %r)
        from mola_fe_lidar_b200 import lidar_odometry as lom
        fake = ctypes.CDLL(%r)
        for f in ("fake_align_calls", "fake_batch_calls", "fake_voxel_calls"):
            getattr(fake, f).restype = ctypes.c_ulong
        def scan(x, y=0.0, z=0.0, n=64):
            a = np.zeros((n, 3), np.float32); a[:, 0] = x; a[:, 1] = y; a[:, 2] = z
            a[1:] += np.random.default_rng(0).normal(0, 1, (n - 1, 3)).astype(np.float32)
            return a
        out = {}
    """ % (ROOT, fake_lib)) + textwrap.dedent(body) + "\nprint('RESULT ' + json.dumps(out))\n"
    env = dict(os.environ, LD_PRELOAD=fake_lib)
    p = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stderr[-3000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")][-1]
    return json.loads(line[len("RESULT "):])


def test_static_libstdcxx_is_not_linked_in():
    """A second, statically linked copy of libstdc++ inside our shared objects
    crashes iostream formatting inside python (the image's CXX wrapper links
    that way); the Makefiles must use the distribution compiler."""
    for lib in ("libmola_fe_lidar_b200.so", "libb200icp.so"):
        path = os.path.join(ROOT, "mola-fe-lidar_b200", "lib", lib)
        syms = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True).stdout
        assert "_ZNSo9_M_insertIdEERSoT_" not in syms, f"{lib} carries a static libstdc++"


def test_params_dump_and_keyframe_rule(fake_lib):
    r = run_child(fake_lib, """
        lo = lom.LidarOdometry(yaml_text=lom.system_yaml())
        out["params"] = lo.params()
        # 1 m forward per 0.1 s scan: KF rule = goodness > 0.5 and dist > 3 m (kitti-default.yaml:8,12)
        for i in range(9):
            lo.onNewObservation(scan(float(i)), 0.1 * i, sync=True)
        s = lo.state()
        out["n_keyframes"] = int(s["n_keyframes"]); out["n_factors"] = int(s["n_factors"])
        out["n_icp"] = int(s["n_icp"]); out["twist"] = [float(v) for v in s["last_twist"]]
        out["accum"] = [float(v) for v in s["accum_since_last_kf"]]
        out["factors"] = [(int(a), int(b), [float(v) for v in p]) for a, b, p in lo.factors()]
        out["align_calls"] = int(fake.fake_align_calls())
        out["profile"] = sorted(lo.profile().keys())
        lo.close()
    """)
    p = r["params"]
    assert float(p["min_time_between_scans"]) == 0.01 and float(p["min_dist_xyz_between_keyframes"]) == 3
    assert int(p["loop_closure_montecarlo_samples"]) == 10 and int(p["icp[2].maxIterations"]) == 100
    # scan 0 -> KF without ICP (cpp:250-257); then a KF each time the accumulated
    # motion EXCEEDS 3 m, i.e. after 4 scans of 1 m: scans 4 and 8
    assert r["n_icp"] == 8 and r["align_calls"] >= 8
    assert r["n_keyframes"] == 3
    assert r["n_factors"] >= 2
    first_two = [f for f in r["factors"] if f[1] == f[0] + 1][:2]
    for f in first_two:
        assert abs(f[2][0] - 4.0) < 1e-9 and abs(f[2][1]) < 1e-12
    # twist = pose / dt (cpp:305-308): 1 m / 0.1 s
    assert abs(r["twist"][0] - 10.0) < 1e-6
    assert abs(r["accum"][0]) < 1e-9  # reset at the last key-frame (cpp:472-474)
    # the reference's profiler section names
    for name in ("doProcessNewObservation", "doProcessNewObservation.3.icp_latest",
                 "doProcessNewObservation.1.filter_pointclouds", "run_one_icp"):
        assert name in r["profile"], name


def test_time_gate_label_filter_and_low_goodness(fake_lib):
    r = run_child(fake_lib, """
        lo = lom.LidarOdometry(yaml_text=lom.system_yaml())
        lo.onNewObservation(scan(0.0), 0.0, sync=True)
        lo.onNewObservation(scan(1.0), 0.005, sync=True)            # < min_time_between_scans: gated (cpp:202-212)
        lo.onNewObservation(scan(1.0), 0.1, label="other", sync=True)  # not my sensor (cpp:169)
        s = lo.state(); out["after_gate"] = [int(s["n_processed"]), int(s["n_icp"]), int(s["n_dropped"])]
        # z0 < 0 makes the fake report quality 0.1 < min_icp_goodness: no KF however far we move
        for i in range(1, 8):
            lo.onNewObservation(scan(2.0 * i, z=-1.0), 0.1 * i, sync=True)
        s = lo.state(); out["n_keyframes"] = int(s["n_keyframes"]); out["goodness"] = float(s["last_icp_goodness"])
        out["accum_x"] = float(s["accum_since_last_kf"][0])
        lo.reset()
        s = lo.state(); out["after_reset"] = [int(s["n_processed"]), int(s["n_keyframes"]), float(s["last_obs_tim"])]
        lo.close()
    """)
    assert r["after_gate"][0] == 1 and r["after_gate"][1] == 0 and r["after_gate"][2] >= 1
    assert r["n_keyframes"] == 1 and abs(r["goodness"] - 0.1) < 1e-12
    assert r["accum_x"] > 10.0  # keeps accumulating while no KF is accepted
    assert r["after_reset"][0] == 0


def test_voxel_stage_from_pointcloud_filter_block(fake_lib):
    r = run_child(fake_lib, """
        lo = lom.LidarOdometry(yaml_text=lom.system_yaml(voxel_resolution=1.0))
        for i in range(3):
            lo.onNewObservation(scan(float(i), n=100), 0.1 * i, sync=True)
        out["voxel_calls"] = int(fake.fake_voxel_calls())
        out["last_points_size"] = int(lo.state()["last_points_size"])
        out["res"] = lo.params()["voxel_decimation_resolution"]
        lo.close()
    """)
    assert r["voxel_calls"] == 3 and r["last_points_size"] == 50 and float(r["res"]) == 1.0


def test_edges_planes_stage_from_pointcloud_filter_block(fake_lib):
    """`pointcloud_filter: - class_name: ...FilterEdgesPlanes` (the class the reference's stale keys name,
    kitti-default.yaml:21-32): parsed with the shipped defaults, the configured layer is what gets registered
    (the double returns layers of n/4, n/3 and n/2 points), unknown layers / classes are refused."""
    r = run_child(fake_lib, """
        def block(cls, layer):
            return ("  pointcloud_filter:\\n"
                    "    - class_name: " + cls + "\\n"
                    "      params:\\n"
                    "        voxel_filter_resolution: 0.5\\n"
                    "        voxel_filter_decimation: 2\\n"
                    "        b200_register_layer: " + layer + "\\n")
        sizes = {}
        for cls, layer in (("mp2p_icp_filters::FilterEdgesPlanes", "edges"),
                           ("mola::lidar_segmentation::FilterEdgesPlanes", "planes"),
                           ("mp2p_icp_filters::FilterEdgesPlanes", "full_decim")):
            lo = lom.LidarOdometry(yaml_text=lom.system_yaml(extra=block(cls, layer)))
            for i in range(2):
                lo.onNewObservation(scan(float(i), n=120), 0.1 * i, sync=True)
            sizes[layer] = int(lo.state()["last_points_size"])
            out["n_icp_" + layer] = int(lo.state()["n_icp"])
            lo.close()
        out["sizes"] = sizes
        errs = []
        for cls, layer in (("mp2p_icp_filters::FilterEdgesPlanes", "corners"), ("mp2p_icp_filters::FilterNoSuch", "edges")):
            try:
                lom.LidarOdometry(yaml_text=lom.system_yaml(extra=block(cls, layer)))
                errs.append("")
            except Exception as e:
                errs.append(str(e))
        out["errs"] = errs
    """)
    assert r["sizes"] == {"edges": 30, "planes": 40, "full_decim": 60}
    assert r["n_icp_edges"] == 1 and r["n_icp_planes"] == 1 and r["n_icp_full_decim"] == 1
    assert "b200_register_layer" in r["errs"][0] and "not registered" in r["errs"][1]


def test_async_queue_and_extra_edges_use_batch_api(fake_lib):
    """onNewObservation is asynchronous (1-thread pool, cpp:183-184); the extra
    nearby / loop-closure edges run on the second pool (cpp:711-729) and the
    Monte-Carlo loop (cpp:775-787) goes through b200icp_align_batch."""
    r = run_child(fake_lib, """
        lo = lom.LidarOdometry(yaml_text=lom.system_yaml())
        # out 80 m along x, back along y = 6 m: key-frames of the return leg come
        # within [min_dist_to_matching, max_dist_to_matching] of the outbound ones
        k = 0
        for i in range(41):
            lo.onNewObservation(scan(2.0 * i), 0.1 * k); k += 1; lo.wait_idle()
        for i in range(40, -1, -1):
            lo.onNewObservation(scan(2.0 * i, y=6.0), 0.1 * k); k += 1; lo.wait_idle()
        lo.wait_idle()
        s = lo.state()
        out["n_keyframes"] = int(s["n_keyframes"]); out["n_checked_pairs"] = int(s["n_checked_pairs"])
        out["n_factors"] = int(s["n_factors"]); out["n_graph_edges"] = int(s["n_graph_edges"])
        out["batch_calls"] = int(fake.fake_batch_calls()); out["align_calls"] = int(fake.fake_align_calls())
        out["profile"] = sorted(lo.profile().keys())
        lo.close()
    """)
    assert r["n_keyframes"] >= 20
    assert r["n_checked_pairs"] >= 1, r
    assert r["n_factors"] >= r["n_keyframes"] - 1
    assert "onNewObservation" in r["profile"] and "delay_onNewObs_to_process" in r["profile"]


def test_additive_key_disables_extra_edge_checks(fake_lib):
    """b200_extra_edge_checks: false (additive key, used by bench.py's e2e leg)
    skips checkForNearbyKFs (cpp:493-508): key-frames and odometry factors are
    unchanged, no pair is ever sent to the second pool."""
    body = """
        lo = lom.LidarOdometry(yaml_text=lom.system_yaml(extra=%r))
        k = 0
        for i in range(21):
            lo.onNewObservation(scan(2.0 * i), 0.1 * k, sync=True); k += 1
        for i in range(20, -1, -1):
            lo.onNewObservation(scan(2.0 * i, y=6.0), 0.1 * k, sync=True); k += 1
        lo.wait_idle()
        s = lo.state()
        out["n_keyframes"] = int(s["n_keyframes"]); out["n_checked_pairs"] = int(s["n_checked_pairs"])
        out["n_icp"] = int(s["n_icp"]); out["n_processed"] = int(s["n_processed"])
        out["profile"] = sorted(lo.profile().keys())
        lo.close()
    """
    on = run_child(fake_lib, body % "")
    off = run_child(fake_lib, body % "  b200_extra_edge_checks: false\n")
    assert on["n_keyframes"] == off["n_keyframes"] >= 10
    assert on["n_checked_pairs"] >= 1 and off["n_checked_pairs"] == 0
    assert off["n_icp"] == off["n_processed"] - 1  # exactly one registration per scan after the first
    assert "doProcessNewObservation.6.checkForNearbyKFs" not in off["profile"]
    assert "doProcessNewObservation.0.upload_and_index" in off["profile"]


def test_keyframe_store_spills_and_reloads_without_changing_results(fake_lib):
    """Key-frame cloud store: with a budget (additive key
    b200_kf_store_budget_mb) the least recently used key-frame clouds are
    spilled to host memory and re-uploaded when an extra-edge registration needs
    them; key-frames, factors and poses are those of the unlimited run."""
    body = """
        fake.fake_download_calls.restype = ctypes.c_ulong
        lo = lom.LidarOdometry(yaml_text=lom.system_yaml(extra=%r))
        k = 0
        for i in range(31):
            lo.onNewObservation(scan(2.0 * i, n=1000), 0.1 * k); k += 1; lo.wait_idle()
        for i in range(30, -1, -1):
            lo.onNewObservation(scan(2.0 * i, y=6.0, n=1000), 0.1 * k); k += 1; lo.wait_idle()
        lo.wait_idle()
        s = lo.state()
        out["n_keyframes"] = int(s["n_keyframes"]); out["n_factors"] = int(s["n_factors"])
        out["n_checked_pairs"] = int(s["n_checked_pairs"]); out["n_icp"] = int(s["n_icp"])
        out["spills"] = int(s["n_kf_spills"]); out["reloads"] = int(s["n_kf_reloads"])
        out["downloads"] = int(fake.fake_download_calls())
        out["factors"] = sorted((int(a), int(b), [round(float(v), 9) for v in p]) for a, b, p in lo.factors())
        lo.close()
    """
    # the test double charges 100 bytes per point: 1000-point clouds = 100 kB each, budget = 3 clouds
    free = run_child(fake_lib, body % "")
    tight = run_child(fake_lib, body % "  b200_kf_store_budget_mb: 0.3\n")
    assert free["spills"] == 0 and free["reloads"] == 0 and free["downloads"] == 0
    assert tight["spills"] >= free["n_keyframes"] - 4 and tight["downloads"] == tight["spills"]
    assert tight["reloads"] >= 1  # the return leg registers against spilled key-frames of the outbound leg
    for key in ("n_keyframes", "n_factors", "n_checked_pairs", "n_icp", "factors"):
        assert tight[key] == free[key], key


def test_cpp_example_builds_and_runs_against_the_test_double(fake_lib, tmp_path):
    """examples/odometry_cpp.cpp: the module driven from C++ (initialize(Yaml) +
    onNewObservation, as mola-launcher drives the reference) links against the
    two shared libraries and runs end to end with the device test double."""
    exe = str(tmp_path / "odometry_cpp")
    lib_dir = os.path.join(ROOT, "mola-fe-lidar_b200", "lib")
    subprocess.check_call(["/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++", "-std=c++17", "-O1",
                           "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "mola-fe-lidar_b200", "host"),
                           os.path.join(ROOT, "examples", "odometry_cpp.cpp"), "-L", lib_dir,
                           "-lmola_fe_lidar_b200", "-lb200icp", f"-Wl,-rpath,{lib_dir}", "-lpthread", "-o", exe])
    p = subprocess.run([exe, os.path.join(ROOT, "mola-fe-lidar_b200")], env=dict(os.environ, LD_PRELOAD=fake_lib),
                       capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stderr[-2000:]
    assert "scan 7: processed 8, registrations" in p.stdout and "factors:" in p.stdout
