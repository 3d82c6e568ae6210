"""FilterEdgesPlanes on the device (edges_planes.cu; the filter class the reference's parameter files name,
kitti-default.yaml:21-32) against the oracle's restatement (A.13): per-point layer flags bit-exact, the three
output clouds = the flagged points in ascending original index, and the module registering one of the layers."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _check(icp, oracle, pts, **kw):
    pts = np.ascontiguousarray(pts, dtype=np.float32)
    g = icp.upload_raw(pts) if hasattr(icp, "upload_raw") else icp.upload(pts)
    layers, flags, nv = icp.filter_edges_planes(g, **kw)
    oflags, onv = oracle.filter_edges_planes(pts, **kw)
    assert np.array_equal(flags, oflags), f"{(flags != oflags).sum()} of {len(flags)} layer flags differ"
    assert nv == onv
    for bit, cloud in enumerate(layers):
        want = pts[(oflags >> bit) & 1 == 1]
        got = cloud.download()
        assert got.shape == want.shape
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    out = [len(c) for c in layers]
    for c in layers:
        c.free()
    g.free()
    return out


def test_lidar_scan_shipped_parameters(icp, oracle):
    from mola_fe_lidar_b200 import scene
    scans, _ = scene.make_sequence(2, seed=3)
    for s in scans:
        n_edges, n_planes, n_full = _check(icp, oracle, s)
        assert n_planes > 0 and n_full > 0


@pytest.mark.parametrize("res,vd,fd", [(0.5, 1, 1), (2.0, 3, 7), (0.25, 2, 20)])
def test_parameter_sweep(icp, oracle, rng, res, vd, fd):
    from mola_fe_lidar_b200 import scene
    pts = scene.make_pair_c1(seed=2, n=20000, sigma=0.01)[0]
    _check(icp, oracle, pts, voxel_filter_resolution=res, voxel_filter_decimation=vd, full_pointcloud_decimation=fd)


def test_shapes_and_degenerate_inputs(icp, oracle, rng):
    wall = np.c_[rng.uniform(0.01, 0.99, 300), 0.5 + rng.normal(0, 1e-3, 300), rng.uniform(0.01, 0.99, 300)]
    ground = np.c_[rng.uniform(0.01, 0.99, 300) + 2, rng.uniform(0.01, 0.99, 300), 0.5 + rng.normal(0, 1e-3, 300)]
    blob = rng.uniform(0.01, 0.99, (300, 3)) + np.array([4.0, 0, 0])
    line = np.c_[rng.uniform(0.01, 0.99, 300) + 6, 0.5 + rng.normal(0, 1e-3, 300), 0.5 + rng.normal(0, 1e-3, 300)]
    few = rng.uniform(0.01, 0.99, (4, 3)) + np.array([8.0, 0, 0])            # below min_points_per_voxel
    dup = np.repeat(np.array([[10.5, 0.5, 0.5]]), 50, axis=0)                # zero covariance
    bad = np.array([[np.nan, 0, 0], [np.inf, 1, 1], [0.5, 0.5, 3e7]])        # NaN / inf / outside the key range
    pts = np.concatenate([wall, ground, blob, line, few, dup, bad]).astype(np.float32)
    pts = pts[rng.permutation(len(pts))]
    n_edges, n_planes, _ = _check(icp, oracle, pts, voxel_filter_decimation=1, full_pointcloud_decimation=1)
    assert n_planes == 300 and n_edges >= 300     # the wall is a plane, the ground is dropped, the blob is "edges"
    _check(icp, oracle, pts[:0])                   # empty cloud


def test_module_registers_the_configured_layer(oracle):
    """pointcloud_filter: FilterEdgesPlanes in the module (additive block, as for the voxel filter): every scan is
    registered as the configured layer of the filter, so pose / quality / iterations per scan equal the oracle's ICP
    run on the oracle's layers with the module's velocity-model guesses (cpp:272-275)."""
    from mola_fe_lidar_b200 import lidar_odometry, scene
    scans, _ = scene.make_sequence(4, seed=1)
    extra = ("  pointcloud_filter:\n"
             "    - class_name: mp2p_icp_filters::FilterEdgesPlanes\n"
             "      params:\n"
             "        voxel_filter_resolution: 1.0\n"
             "        voxel_filter_decimation: 1\n"
             "        full_pointcloud_decimation: 4\n"
             "        b200_register_layer: full_decim\n"
             "  b200_extra_edge_checks: false\n")
    lo = lidar_odometry.LidarOdometry(yaml_text=lidar_odometry.system_yaml(extra=extra))
    layers = []
    for s in scans:
        f, _ = oracle.filter_edges_planes(s, voxel_filter_decimation=1, full_pointcloud_decimation=4)
        layers.append(oracle.Cloud(np.ascontiguousarray(s[(f >> 2) & 1 == 1])))
    prm = oracle.default_params()
    twist, dt = np.zeros(4), 0.1
    for i, s in enumerate(scans):
        lo.onNewObservation(s, dt * i, sync=True)
        if i == 0:
            continue
        st = lo.state()
        guess = np.array([twist[0] * dt, twist[1] * dt, twist[2] * dt, twist[3] * dt, 0, 0])
        r = oracle.icp_align(layers[i - 1], layers[i], guess, prm, kdtree=True)
        assert st["n_icp"] == i
        assert np.abs(st["last_icp_pose"][:3] - r["pose"][:3]).max() < 1e-5
        assert np.abs(st["last_icp_pose"][3:] - r["pose"][3:]).max() < 1e-6
        assert st["last_icp_goodness"] == r["quality"] and st["last_icp_iterations"] == r["n_iterations"]
        twist = np.array([r["pose"][0] / dt, r["pose"][1] / dt, r["pose"][2] / dt, r["pose"][3] / dt])
    lo.close()
    with pytest.raises(Exception, match="b200_register_layer"):
        lidar_odometry.LidarOdometry(yaml_text=lidar_odometry.system_yaml(
            extra=extra.replace("full_decim", "corners")))


def test_matches_independent_golden(icp):
    """The device against tests/golden/edges_planes_c1.npz (numpy eigh on np.unique voxels, unambiguous voxels only)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "edges_planes_c1.npz"))
    cloud = icp.upload_raw(np.ascontiguousarray(g["pts"]))
    layers, f, _ = icp.filter_edges_planes(cloud, voxel_filter_decimation=1, full_pointcloud_decimation=1)
    cls_of_point = np.where(f & 1, 1, np.where(f & 2, 2, 0))
    want = g["voxel_class"][g["voxel_of_point"]]
    sure = g["voxel_sure"][g["voxel_of_point"]]
    assert np.array_equal(cls_of_point[sure], want[sure])
    for c in layers:
        c.free()
    cloud.free()
