"""CPU tests (no GPU): YAML loading of the shipped parameter files through the
C ABI, exported symbols of both libraries vs include/*.h, loud failure without
a device, synthetic scene generator, bench helpers."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_PARAMS = "/root/reference/params"


def test_build_entry_point():
    import __graft_entry__ as g
    g.build()
    from mola_fe_lidar_b200 import capi, lidar_odometry
    assert os.path.exists(capi.LIB_PATH) and os.path.exists(lidar_odometry.LIB_PATH)


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(b200(?:icp|lo)_[a-z0-9_]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported(capi):
    from mola_fe_lidar_b200 import lidar_odometry
    L = capi.lib()
    decl = _declared("b200icp.h")
    assert len(decl) >= 20
    for s in decl:
        assert hasattr(L, s), f"{s} declared in include/b200icp.h but not exported"
    assert sorted(capi.EXPORTS) == decl
    L2 = lidar_odometry.lib()
    decl2 = _declared("b200_lidar_odometry.h")
    for s in decl2:
        assert hasattr(L2, s), f"{s} declared in include/b200_lidar_odometry.h but not exported"
    assert sorted(lidar_odometry.EXPORTS) == decl2


def test_no_cpu_fallback_without_device(capi):
    if capi.device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(capi.B200IcpError, match="(?i)cuda"):
        capi.ICP(capi.default_params())
    from mola_fe_lidar_b200 import lidar_odometry
    with pytest.raises(capi.B200IcpError):
        lidar_odometry.LidarOdometry()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "mola-fe-lidar_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle_api" not in txt and "icp_oracle" not in txt and "liboracle" not in txt, f


def test_shipped_icp_yaml_parses(capi):
    from mola_fe_lidar_b200 import lidar_odometry
    for name in ("icp-settings-regular.yaml", "icp-settings-loop-closure.yaml"):
        p = capi.params_from_yaml(open(os.path.join(lidar_odometry.PARAMS_DIR, name)).read())
        assert p.max_iterations == 100 and p.min_abs_step_trans == 5e-5 and p.min_abs_step_rot == 1e-5
        assert p.use_scale_outlier_detector == 1 and p.scale_outlier_threshold == 1.1
        assert p.use_robust_kernel == 0 and abs(p.robust_kernel_param - np.deg2rad(0.1)) < 1e-15
        assert p.robust_kernel_scale == 400.0
        assert p.solver_kind == capi.SOLVER_GAUSS_NEWTON and p.solver_max_iterations == 20
        assert p.matcher_kind == capi.MATCHER_POINT2PLANE and p.distance_threshold == 0.70
        assert p.plane_eigen_threshold == 0.07 and p.knn == 6
        assert p.run_from_iteration == 0 and p.run_up_to_iteration == 0
        assert p.quality_threshold_distance == 0.10


@pytest.mark.skipif(not os.path.isdir(REF_PARAMS), reason="reference tree not mounted")
def test_reference_yaml_files_parse_unchanged(capi):
    """The reference's own files, byte for byte, give the same parameters as ours."""
    from mola_fe_lidar_b200 import lidar_odometry
    for name in ("icp-settings-regular.yaml", "icp-settings-loop-closure.yaml"):
        ref = capi.params_from_yaml(open(os.path.join(REF_PARAMS, name)).read())
        ours = capi.params_from_yaml(open(os.path.join(lidar_odometry.PARAMS_DIR, name)).read())
        assert ref.as_dict() == ours.as_dict()


def test_yaml_errors_name_the_class(capi):
    from mola_fe_lidar_b200 import lidar_odometry
    txt = open(os.path.join(lidar_odometry.PARAMS_DIR, "icp-settings-regular.yaml")).read()
    with pytest.raises(capi.B200IcpError, match="foo::Bar"):
        capi.params_from_yaml(txt.replace("mp2p_icp::ICP", "foo::Bar"))
    with pytest.raises(capi.B200IcpError, match="Matcher_Nope"):
        capi.params_from_yaml(txt.replace("Matcher_Point2Plane", "Matcher_Nope"))
    with pytest.raises(capi.B200IcpError, match="solvers"):
        capi.params_from_yaml(txt.replace("solvers:", "solverz:"))
    with pytest.raises(capi.B200IcpError, match="icp_class"):
        capi.params_from_yaml(txt.replace("icp_class:", "icp_klass:"))
    p = capi.params_from_yaml(txt.replace("Solver_GaussNewton", "Solver_Horn")
                              .replace("Matcher_Point2Plane", "Matcher_Points_DistanceThreshold"))
    assert p.solver_kind == capi.SOLVER_HORN and p.matcher_kind == capi.MATCHER_POINTS_DISTANCE


def test_scene_generator_shapes_and_determinism():
    from mola_fe_lidar_b200 import scene
    s1, p1 = scene.make_sequence(2, seed=5)
    s2, _ = scene.make_sequence(2, seed=5)
    assert s1[0].shape == (120000, 3) and s1[0].dtype == np.float32
    assert np.array_equal(s1[1], s2[1])
    rel = scene.relative_pose6(p1[0], p1[1])
    assert abs(rel[0] - 1.0) < 1e-9 and abs(rel[3] - 0.005) < 1e-12
    A, B, pose = scene.make_pair_c1(seed=1, n=500)
    T = scene.pose_matrix(*pose)
    assert np.allclose(B.astype(np.float64) @ T[:3, :3].T + T[:3, 3], A, atol=1e-4)


def test_bench_helpers():
    sys.path.insert(0, ROOT)
    import bench
    idx = [bench.scan_index(i) for i in range(40)]
    assert all(abs(a - b) == 1 for a, b in zip(idx, idx[1:])) and min(idx) == 0 and max(idx) == bench.N_SCANS - 1
    peak, src = bench.load_peaks()
    assert peak > 1000 and src in ("measured", "fallback")


def test_world_size_2_gloo_sharding(tmp_path):
    """N>1 path on CPU: two ranks shard independent registration units and
    reduce (max time, sum of units) exactly like bench.py does over NCCL."""
    code = r'''
import os, sys, torch, torch.distributed as dist
dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
units = list(range(11))[r::w]                      # round-robin pairs -> ranks, no data-path collective
t = torch.tensor([100.0 + 10 * r, 50.0], dtype=torch.float64)
u = torch.tensor([float(len(units)), 1.0], dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX); dist.all_reduce(u, op=dist.ReduceOp.SUM)
assert float(t[0]) == 110.0 and float(u[0]) == 11.0 and float(u[1]) == w
gathered = [None] * w
dist.all_gather_object(gathered, units)
assert sorted(sum(gathered, [])) == list(range(11))
# packed (d2 bits << 32 | idx) keys: MIN over ranks = global arg-min with the lowest-index tie rule
import numpy as np
d2 = np.float32([0.25, 0.5][r]); idx = [7, 3][r]
key = torch.tensor([(int(np.float32(d2).view(np.uint32)) << 32) | idx], dtype=torch.int64)
dist.all_reduce(key, op=dist.ReduceOp.MIN)
assert int(key[0]) & 0xFFFFFFFF == 7
dist.destroy_process_group()
'''
    script = tmp_path / "gloo_ranks.py"
    script.write_text(code)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
