"""Voxel decimation on the device vs the oracle (row F, Appendix A.11):
kept indices and output coordinates bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("res,avg", [(1.0, False), (0.5, False), (0.25, True), (1.0, True)])
def test_voxel_scan(icp, oracle, res, avg):
    from mola_fe_lidar_b200 import scene
    scans, _ = scene.make_sequence(1, seed=1)
    pts = scans[0]
    g = icp.upload(pts)
    out, keep = icp.voxel_decimate(g, res, use_average=avg, want_indices=True)
    okeep, oxyz = oracle.voxel_decimate(pts, res, use_average=avg)
    assert np.array_equal(keep, okeep)
    xyz = out.download()
    assert np.array_equal(xyz.view(np.uint32), oxyz.view(np.uint32))
    assert (np.diff(keep.astype(np.int64)) > 0).all()
    g.free(), out.free()


def test_voxel_edge_cases(icp, oracle, rng):
    pts = rng.uniform(-3, 3, size=(5000, 3)).astype(np.float32)
    pts[10] = np.nan
    pts[20:40] = pts[19]                      # duplicates collapse into one voxel
    pts[100] = [-0.0, 0.0, -1e-30]            # negative side of zero: floor -> -1
    g = icp.upload(pts)
    out, keep = icp.voxel_decimate(g, 0.4, want_indices=True)
    okeep, oxyz = oracle.voxel_decimate(pts, 0.4)
    assert np.array_equal(keep, okeep) and 10 not in keep
    assert np.array_equal(out.download().view(np.uint32), oxyz.view(np.uint32))
    # idempotence: decimating the decimated cloud keeps everything
    out2 = icp.voxel_decimate(out, 0.4)
    assert len(out2) == len(out)
    empty = icp.upload(np.zeros((0, 3), dtype=np.float32))
    assert len(icp.voxel_decimate(empty, 1.0)) == 0
    # the decimated cloud is searchable
    idx, _ = icp.knn(out, out, 1, 0.7)
    assert np.array_equal(idx[:, 0], np.arange(len(out), dtype=np.uint32))


def test_raw_upload_feeds_the_filter_only(icp, oracle, capi):
    """apply_generators -> apply_filter_pipeline (LidarOdometry.cpp:215-224): the raw scan needs no search index
    when a filter stage follows; its output is the same cloud, and a coordinates-only cloud is refused as a
    side of a search."""
    from mola_fe_lidar_b200 import scene
    scans, _ = scene.make_sequence(1, seed=2)
    raw = icp.upload_raw(scans[0])
    out, keep = icp.voxel_decimate(raw, 1.0, want_indices=True)
    okeep, oxyz = oracle.voxel_decimate(scans[0], 1.0)
    assert np.array_equal(keep, okeep)
    assert np.array_equal(out.download().view(np.uint32), oxyz.view(np.uint32))
    with pytest.raises(capi.B200IcpError):
        icp.knn(raw, out, 1, 0.7)
    raw.free(), out.free()


@pytest.mark.parametrize("res", [1.0, 0.5])
def test_decimated_clouds_search_exact_with_coarse_cells(icp, oracle, res):
    """A decimated cloud is indexed with cells that follow the voxel size (one point per voxel): neighbour lists
    still equal the oracle's exact search, both as reference and as query side."""
    from mola_fe_lidar_b200 import scene
    scans, poses = scene.make_sequence(2, seed=3)
    a = icp.voxel_decimate(icp.upload_raw(scans[0]), res)
    b = icp.voxel_decimate(icp.upload_raw(scans[1]), res)
    A, B = a.download(), b.download()
    idx, d2 = icp.knn(a, b, 6, 0.7)
    oidx, od2 = oracle.knn(oracle.Cloud(A), B, 6, 0.49, kdtree=True)
    assert np.array_equal(idx, oidx)
    assert np.array_equal(d2.view(np.uint32), od2.view(np.uint32))
    # the decimated cloud against a raw (finely indexed) one and the other way round
    raw = icp.upload(scans[0])
    idx2, _ = icp.knn(raw, b, 6, 0.7)
    oidx2, _ = oracle.knn(oracle.Cloud(scans[0]), B, 6, 0.49, kdtree=True)
    assert np.array_equal(idx2, oidx2)
    idx3, _ = icp.knn(a, raw, 1, 0.7)
    oidx3, _ = oracle.knn(oracle.Cloud(A), scans[0], 1, 0.49, kdtree=True)
    assert np.array_equal(idx3, oidx3)
