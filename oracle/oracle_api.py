"""ctypes view of oracle/liboracle.so -- TEST INFRASTRUCTURE ONLY.

Imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs (as the checker or the timed CPU baseline).  The
product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")

INVALID = 0xFFFFFFFF
TERM = {0: "Undefined", 1: "NoPairings", 2: "SolverError", 3: "MaxIterations", 4: "Stalled"}


class Params(C.Structure):
    _fields_ = [
        ("max_iterations", C.c_uint32),
        ("min_abs_step_trans", C.c_double),
        ("min_abs_step_rot", C.c_double),
        ("use_scale_outlier_detector", C.c_int32),
        ("scale_outlier_threshold", C.c_double),
        ("use_robust_kernel", C.c_int32),
        ("robust_kernel_param", C.c_double),
        ("robust_kernel_scale", C.c_double),
        ("solver_kind", C.c_int32),
        ("solver_max_iterations", C.c_uint32),
        ("gn_min_delta", C.c_double),
        ("matcher_kind", C.c_int32),
        ("distance_threshold", C.c_double),
        ("plane_eigen_threshold", C.c_double),
        ("knn", C.c_uint32),
        ("min_plane_points", C.c_uint32),
        ("run_from_iteration", C.c_uint32),
        ("run_up_to_iteration", C.c_uint32),
        ("quality_threshold_distance", C.c_double),
        ("cov_fd_step", C.c_double),
    ]


class Result(C.Structure):
    _fields_ = [
        ("pose", C.c_double * 6),
        ("R", C.c_double * 9),
        ("t", C.c_double * 3),
        ("cov", C.c_double * 36),
        ("quality", C.c_double),
        ("n_iterations", C.c_uint32),
        ("termination_reason", C.c_uint32),
        ("n_pairings", C.c_uint32),
        ("cov_singular", C.c_uint32),
    ]


def build(force=False):
    src = os.path.join(_HERE, "icp_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB):
        build()
    L = C.CDLL(_LIB)
    fp = C.POINTER(C.c_float)
    dp = C.POINTER(C.c_double)
    up = C.POINTER(C.c_uint32)
    vp = C.c_void_p
    L.orc_default_params.argtypes = [C.POINTER(Params)]
    L.orc_cloud_create.restype = vp
    L.orc_cloud_create.argtypes = [fp, fp, fp, C.c_size_t]
    L.orc_cloud_free.argtypes = [vp]
    L.orc_cloud_size.restype = C.c_size_t
    L.orc_cloud_size.argtypes = [vp]
    for name in ("orc_knn_brute", "orc_knn_kdtree"):
        f = getattr(L, name)
        f.argtypes = [vp, fp, fp, fp, C.c_size_t, C.c_uint32, C.c_float, up, fp]
        f.restype = None
    L.orc_pose_to_Rt.argtypes = [dp, dp, dp]
    L.orc_Rt_to_pose.argtypes = [dp, dp, dp]
    L.orc_transform_points.argtypes = [dp, dp, fp, fp, fp, C.c_size_t, fp, fp, fp]
    L.orc_se3_exp.argtypes = [dp, dp, dp]
    L.orc_se3_log.argtypes = [dp, dp, dp]
    L.orc_eig3_sym.argtypes = [dp, dp, dp]
    L.orc_qr_solve6.argtypes = [dp, dp, dp]
    L.orc_qr_solve6.restype = C.c_int
    L.orc_match_point2plane.argtypes = [vp, vp, dp, dp, C.POINTER(Params), C.c_int,
                                        C.POINTER(C.c_uint8), up, up, dp, dp]
    L.orc_match_point2plane.restype = C.c_size_t
    L.orc_match_points.argtypes = [vp, vp, dp, dp, C.c_double, C.c_int, up, fp]
    L.orc_match_points.restype = C.c_size_t
    L.orc_gn_point2plane.argtypes = [dp, dp, dp, C.c_size_t, C.c_uint32, C.c_double, dp, dp]
    L.orc_gn_point2plane.restype = C.c_int
    L.orc_gn_point2point.argtypes = [dp, dp, C.c_size_t, C.c_uint32, C.c_double, dp, dp]
    L.orc_gn_point2point.restype = C.c_int
    L.orc_horn.argtypes = [dp, dp, C.c_size_t, C.POINTER(Params), dp, dp, dp]
    L.orc_horn.restype = C.c_size_t
    L.orc_quality_paired_ratio.argtypes = [vp, vp, dp, dp, C.c_double, C.c_int]
    L.orc_quality_paired_ratio.restype = C.c_double
    L.orc_icp_align.argtypes = [vp, vp, dp, C.POINTER(Params), C.c_int, C.POINTER(Result)]
    L.orc_icp_align.restype = C.c_int
    L.orc_voxel_decimate.argtypes = [fp, fp, fp, C.c_size_t, C.c_float, C.c_int, up, fp, fp, fp]
    L.orc_voxel_decimate.restype = C.c_size_t
    L.orc_filter_edges_planes.argtypes = [fp, fp, fp, C.c_size_t, C.c_float, C.c_uint32, C.c_uint32, C.c_float,
                                          C.c_float, C.c_float, C.c_float, C.c_uint32, C.POINTER(C.c_uint8)]
    L.orc_filter_edges_planes.restype = C.c_size_t
    _lib = L
    return L


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a, ct):
    return a.ctypes.data_as(C.POINTER(ct))


def default_params(**kw):
    p = Params()
    lib().orc_default_params(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


class Cloud:
    """SoA float cloud (Appendix A.1)."""

    def __init__(self, xyz):
        xyz = np.asarray(xyz, dtype=np.float32).reshape(-1, 3)
        self.x, self.y, self.z = _f32(xyz[:, 0]), _f32(xyz[:, 1]), _f32(xyz[:, 2])
        self.n = len(self.x)
        self.h = lib().orc_cloud_create(_p(self.x, C.c_float), _p(self.y, C.c_float),
                                        _p(self.z, C.c_float), self.n)

    def __del__(self):
        try:
            if self.h:
                lib().orc_cloud_free(self.h)
                self.h = None
        except Exception:
            pass

    @property
    def xyz(self):
        return np.stack([self.x, self.y, self.z], axis=1)


def knn(ref, q_xyz, k, max_d2=np.inf, kdtree=False):
    q = np.asarray(q_xyz, dtype=np.float32).reshape(-1, 3)
    qx, qy, qz = _f32(q[:, 0]), _f32(q[:, 1]), _f32(q[:, 2])
    nq = len(qx)
    idx = np.empty((nq, k), dtype=np.uint32)
    d2 = np.empty((nq, k), dtype=np.float32)
    f = lib().orc_knn_kdtree if kdtree else lib().orc_knn_brute
    f(ref.h, _p(qx, C.c_float), _p(qy, C.c_float), _p(qz, C.c_float), nq, k,
      C.c_float(max_d2), _p(idx, C.c_uint32), _p(d2, C.c_float))
    return idx, d2


def pose_to_Rt(pose):
    pose = _f64(pose)
    R = np.empty(9)
    t = np.empty(3)
    lib().orc_pose_to_Rt(_p(pose, C.c_double), _p(R, C.c_double), _p(t, C.c_double))
    return R.reshape(3, 3), t


def Rt_to_pose(R, t):
    R = _f64(R).reshape(9)
    t = _f64(t)
    pose = np.empty(6)
    lib().orc_Rt_to_pose(_p(R, C.c_double), _p(t, C.c_double), _p(pose, C.c_double))
    return pose


def transform_points(R, t, xyz):
    xyz = np.asarray(xyz, dtype=np.float32).reshape(-1, 3)
    x, y, z = _f32(xyz[:, 0]), _f32(xyz[:, 1]), _f32(xyz[:, 2])
    ox, oy, oz = np.empty_like(x), np.empty_like(y), np.empty_like(z)
    R = _f64(R).reshape(9)
    t = _f64(t)
    lib().orc_transform_points(_p(R, C.c_double), _p(t, C.c_double), _p(x, C.c_float),
                               _p(y, C.c_float), _p(z, C.c_float), len(x), _p(ox, C.c_float),
                               _p(oy, C.c_float), _p(oz, C.c_float))
    return np.stack([ox, oy, oz], axis=1)


def se3_exp(eps):
    eps = _f64(eps)
    R = np.empty(9)
    t = np.empty(3)
    lib().orc_se3_exp(_p(eps, C.c_double), _p(R, C.c_double), _p(t, C.c_double))
    return R.reshape(3, 3), t


def se3_log(R, t):
    R = _f64(R).reshape(9)
    t = _f64(t)
    eps = np.empty(6)
    lib().orc_se3_log(_p(R, C.c_double), _p(t, C.c_double), _p(eps, C.c_double))
    return eps


def eig3(Cm):
    Cm = _f64(Cm).reshape(9)
    ev = np.empty(3)
    V = np.empty(9)
    lib().orc_eig3_sym(_p(Cm, C.c_double), _p(ev, C.c_double), _p(V, C.c_double))
    return ev, V.reshape(3, 3)


def qr_solve6(A, b):
    A = _f64(A).reshape(36)
    b = _f64(b)
    x = np.empty(6)
    rank = lib().orc_qr_solve6(_p(A, C.c_double), _p(b, C.c_double), _p(x, C.c_double))
    return x, rank


def match_point2plane(glob, loc, R, t, params, kdtree=False):
    n, k = loc.n, params.knn
    paired = np.zeros(n, dtype=np.uint8)
    nn_idx = np.empty((n, k), dtype=np.uint32)
    nn_cnt = np.zeros(n, dtype=np.uint32)
    cen = np.zeros((n, 3))
    nor = np.zeros((n, 3))
    R = _f64(R).reshape(9)
    t = _f64(t)
    npair = lib().orc_match_point2plane(glob.h, loc.h, _p(R, C.c_double), _p(t, C.c_double),
                                        C.byref(params), int(kdtree), _p(paired, C.c_uint8),
                                        _p(nn_idx, C.c_uint32), _p(nn_cnt, C.c_uint32),
                                        _p(cen, C.c_double), _p(nor, C.c_double))
    return dict(n=npair, paired=paired, nn_idx=nn_idx, nn_cnt=nn_cnt, centroid=cen, normal=nor)


def match_points(glob, loc, R, t, threshold, kdtree=False):
    nn = np.empty(loc.n, dtype=np.uint32)
    d2 = np.empty(loc.n, dtype=np.float32)
    R = _f64(R).reshape(9)
    t = _f64(t)
    n = lib().orc_match_points(glob.h, loc.h, _p(R, C.c_double), _p(t, C.c_double),
                               threshold, int(kdtree), _p(nn, C.c_uint32), _p(d2, C.c_float))
    return n, nn, d2


def gn_point2plane(P, Cc, Nn, R, t, max_iters=20, min_delta=1e-10):
    P, Cc, Nn = _f64(P), _f64(Cc), _f64(Nn)
    R = _f64(R).reshape(9).copy()
    t = _f64(t).copy()
    it = lib().orc_gn_point2plane(_p(P, C.c_double), _p(Cc, C.c_double), _p(Nn, C.c_double),
                                  len(P), max_iters, min_delta, _p(R, C.c_double),
                                  _p(t, C.c_double))
    return R.reshape(3, 3), t, it


def gn_point2point(P, Q, R, t, max_iters=20, min_delta=1e-10):
    P, Q = _f64(P), _f64(Q)
    R = _f64(R).reshape(9).copy()
    t = _f64(t).copy()
    it = lib().orc_gn_point2point(_p(P, C.c_double), _p(Q, C.c_double), len(P), max_iters,
                                  min_delta, _p(R, C.c_double), _p(t, C.c_double))
    return R.reshape(3, 3), t, it


def horn(P, Q, params, Rprior=None):
    P, Q = _f64(P), _f64(Q)
    Rp = _f64(np.eye(3) if Rprior is None else Rprior).reshape(9)
    R = np.empty(9)
    t = np.empty(3)
    used = lib().orc_horn(_p(P, C.c_double), _p(Q, C.c_double), len(P), C.byref(params),
                          _p(Rp, C.c_double), _p(R, C.c_double), _p(t, C.c_double))
    return R.reshape(3, 3), t, used


def quality(glob, loc, R, t, threshold, kdtree=False):
    R = _f64(R).reshape(9)
    t = _f64(t)
    return lib().orc_quality_paired_ratio(glob.h, loc.h, _p(R, C.c_double), _p(t, C.c_double),
                                          threshold, int(kdtree))


def icp_align(from_global, to_local, guess, params, kdtree=True):
    g = _f64(guess)
    r = Result()
    lib().orc_icp_align(from_global.h, to_local.h, _p(g, C.c_double), C.byref(params),
                        int(kdtree), C.byref(r))
    return dict(pose=np.array(r.pose), R=np.array(r.R).reshape(3, 3), t=np.array(r.t),
                cov=np.array(r.cov).reshape(6, 6), quality=r.quality,
                n_iterations=r.n_iterations, termination_reason=r.termination_reason,
                n_pairings=r.n_pairings, cov_singular=r.cov_singular)


def voxel_decimate(xyz, resolution, use_average=False):
    xyz = np.asarray(xyz, dtype=np.float32).reshape(-1, 3)
    x, y, z = _f32(xyz[:, 0]), _f32(xyz[:, 1]), _f32(xyz[:, 2])
    n = len(x)
    keep = np.empty(max(n, 1), dtype=np.uint32)
    ox, oy, oz = (np.empty(max(n, 1), dtype=np.float32) for _ in range(3))
    m = lib().orc_voxel_decimate(_p(x, C.c_float), _p(y, C.c_float), _p(z, C.c_float), n,
                                 C.c_float(resolution), int(use_average), _p(keep, C.c_uint32),
                                 _p(ox, C.c_float), _p(oy, C.c_float), _p(oz, C.c_float))
    return keep[:m].copy(), np.stack([ox[:m], oy[:m], oz[:m]], axis=1)


EDGES_PLANES_DEFAULTS = dict(voxel_filter_resolution=1.0, full_pointcloud_decimation=10, voxel_filter_decimation=10,
                             voxel_filter_max_e2_e0=30.0, voxel_filter_max_e1_e0=30.0, voxel_filter_min_e2_e0=80.0,
                             voxel_filter_min_e1_e0=80.0, min_points_per_voxel=5)  # params/kitti-default.yaml:23-32


def filter_edges_planes(xyz, **kw):
    """A.13: per-point layer flags (bit 0 edges, bit 1 planes, bit 2 full_decim) and the classified voxel count."""
    p = dict(EDGES_PLANES_DEFAULTS, **kw)
    xyz = np.asarray(xyz, dtype=np.float32).reshape(-1, 3)
    x, y, z = _f32(xyz[:, 0]), _f32(xyz[:, 1]), _f32(xyz[:, 2])
    layer = np.zeros(len(x), dtype=np.uint8)
    nv = lib().orc_filter_edges_planes(_p(x, C.c_float), _p(y, C.c_float), _p(z, C.c_float), len(x),
                                       p["voxel_filter_resolution"], p["full_pointcloud_decimation"],
                                       p["voxel_filter_decimation"], p["voxel_filter_max_e2_e0"],
                                       p["voxel_filter_max_e1_e0"], p["voxel_filter_min_e2_e0"],
                                       p["voxel_filter_min_e1_e0"], p["min_points_per_voxel"],
                                       _p(layer, C.c_uint8))
    return layer, int(nv)
