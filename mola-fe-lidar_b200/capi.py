"""ctypes binding of lib/libb200icp.so -- the C ABI of include/b200icp.h.

The same stub a maintainer of the reference would write around the C ABI (see
INTEGRATION.md for the C++ / mp2p_icp side).  There is NO fallback: if the CUDA
library is missing or no B200 is visible every call raises.
"""
import ctypes as C
import os

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# B200ICP_LIB: developer override (A/B of two builds on one box)
LIB_PATH = os.environ.get("B200ICP_LIB") or os.path.join(PKG_DIR, "lib", "libb200icp.so")

INVALID = 0xFFFFFFFF
TERM = {0: "Undefined", 1: "NoPairings", 2: "SolverError", 3: "MaxIterations", 4: "Stalled"}
SOLVER_GAUSS_NEWTON, SOLVER_HORN = 0, 1
MATCHER_POINT2PLANE, MATCHER_POINTS_DISTANCE = 0, 1


class B200IcpError(RuntimeError):
    pass


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

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class CallParams(C.Structure):
    """b200icp_call_params_t: mp2p_icp::Parameters of one align() call (LidarOdometry.cpp:869-871)."""
    _fields_ = [
        ("max_iterations", C.c_uint32),
        ("min_abs_step_trans", C.c_double),
        ("min_abs_step_rot", C.c_double),
        ("use_scale_outlier_detector", C.c_int32),
        ("scale_outlier_threshold", C.c_double),
        ("use_robust_kernel", C.c_int32),
        ("robust_kernel_param", C.c_double),
        ("robust_kernel_scale", C.c_double),
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

    def as_dict(self):
        return dict(pose=np.array(self.pose), R=np.array(self.R).reshape(3, 3), t=np.array(self.t),
                    cov=np.array(self.cov).reshape(6, 6), quality=self.quality,
                    n_iterations=self.n_iterations, termination_reason=self.termination_reason,
                    n_pairings=self.n_pairings, cov_singular=self.cov_singular)


class Profile(C.Structure):
    _fields_ = [
        ("match_launches", C.c_uint64), ("match_ms", C.c_double), ("match_queries", C.c_uint64),
        ("solve_launches", C.c_uint64), ("solve_ms", C.c_double),
        ("index_builds", C.c_uint64), ("index_ms", C.c_double), ("index_points", C.c_uint64),
        ("knn_launches", C.c_uint64), ("knn_ms", C.c_double), ("knn_queries", C.c_uint64),
        ("voxel_launches", C.c_uint64), ("voxel_ms", C.c_double), ("voxel_points", C.c_uint64),
        ("total_kernel_launches", C.c_uint64),
        ("fit_launches", C.c_uint64), ("fit_ms", C.c_double),
        ("graph_replays", C.c_uint64),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


EXPORTS = [
    "b200icp_last_error", "b200icp_device_count", "b200icp_default_params",
    "b200icp_params_from_yaml", "b200icp_create", "b200icp_create_from_yaml", "b200icp_destroy",
    "b200icp_get_params", "b200icp_device", "b200icp_cloud_upload", "b200icp_cloud_upload_raw",
    "b200icp_cloud_from_device",
    "b200icp_cloud_free", "b200icp_cloud_size", "b200icp_cloud_device_bytes", "b200icp_cloud_download",
    "b200icp_voxel_decimate", "b200icp_edges_planes_defaults", "b200icp_filter_edges_planes", "b200icp_knn", "b200icp_knn_keys_device", "b200icp_merge_keys_device",
    "b200icp_knn_keys_scatter", "b200icp_peer_alloc", "b200icp_peer_free", "b200icp_peer_open",
    "b200icp_peer_close", "b200icp_peer_barrier", "b200icp_knn_keys_exchange", "b200icp_fill_no_key",
    "b200icp_match", "b200icp_align", "b200icp_align_with", "b200icp_call_params_of",
    "b200icp_align_batch", "b200icp_profile_enable", "b200icp_profile_reset",
    "b200icp_profile_get", "b200icp_synchronize",
    "b200icp_comm_unique_id", "b200icp_comm_create", "b200icp_comm_destroy", "b200icp_sharded_map_create",
    "b200icp_sharded_map_destroy", "b200icp_sharded_map_local_size", "b200icp_sharded_knn_keys",
    "b200icp_sharded_align",
]

_lib = None


def lib():
    """Loads the CUDA library; raises when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200IcpError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    fp, dp, up, vp = C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_uint32), C.c_void_p
    L.b200icp_last_error.restype = C.c_char_p
    L.b200icp_device_count.restype = C.c_int
    L.b200icp_default_params.argtypes = [C.POINTER(Params)]
    L.b200icp_params_from_yaml.argtypes = [C.c_char_p, C.POINTER(Params)]
    L.b200icp_create.argtypes = [C.POINTER(Params), C.c_int, C.POINTER(vp)]
    L.b200icp_create_from_yaml.argtypes = [C.c_char_p, C.c_int, C.POINTER(vp)]
    L.b200icp_destroy.argtypes = [vp]
    L.b200icp_destroy.restype = None
    L.b200icp_get_params.argtypes = [vp, C.POINTER(Params)]
    L.b200icp_device.argtypes = [vp]
    L.b200icp_cloud_upload.argtypes = [vp, vp, vp, vp, C.c_size_t, C.c_float, C.POINTER(vp)]
    L.b200icp_cloud_upload_raw.argtypes = [vp, vp, vp, vp, C.c_size_t, C.POINTER(vp)]
    L.b200icp_cloud_from_device.argtypes = [vp, vp, vp, vp, C.c_size_t, C.c_float, C.POINTER(vp)]
    L.b200icp_cloud_free.argtypes = [vp]
    L.b200icp_cloud_free.restype = None
    L.b200icp_cloud_size.argtypes = [vp]
    L.b200icp_cloud_size.restype = C.c_size_t
    L.b200icp_cloud_device_bytes.argtypes = [vp]
    L.b200icp_cloud_device_bytes.restype = C.c_size_t
    L.b200icp_cloud_download.argtypes = [vp, fp, fp, fp]
    L.b200icp_voxel_decimate.argtypes = [vp, vp, C.c_float, C.c_int, C.c_float, C.POINTER(vp), up]
    L.b200icp_edges_planes_defaults.argtypes = [C.POINTER(EdgesPlanesParams)]
    L.b200icp_edges_planes_defaults.restype = None
    L.b200icp_filter_edges_planes.argtypes = [vp, vp, C.POINTER(EdgesPlanesParams), C.c_float, C.POINTER(vp * 3),
                                              C.POINTER(C.c_uint8), C.POINTER(C.c_uint32)]
    L.b200icp_knn.argtypes = [vp, vp, vp, dp, C.c_uint32, C.c_float, up, fp]
    L.b200icp_knn_keys_device.argtypes = [vp, vp, vp, dp, C.c_uint32, C.c_float, vp, vp]
    L.b200icp_merge_keys_device.argtypes = [vp, vp, C.c_uint32, C.c_size_t, C.c_size_t, C.c_uint32, vp]
    L.b200icp_knn_keys_scatter.argtypes = [vp, vp, vp, dp, C.c_uint32, C.c_float, vp, C.POINTER(vp),
                                           C.c_uint32, C.c_uint32, C.c_int]
    L.b200icp_peer_alloc.argtypes = [vp, C.c_size_t, C.POINTER(vp), C.c_char_p]
    L.b200icp_peer_free.argtypes = [vp, vp]
    L.b200icp_peer_open.argtypes = [vp, C.c_char_p, C.POINTER(vp)]
    L.b200icp_peer_close.argtypes = [vp, vp]
    L.b200icp_peer_barrier.argtypes = [vp, C.POINTER(vp), C.c_uint32, C.c_uint32, C.c_uint64]
    L.b200icp_knn_keys_exchange.argtypes = [vp, vp, vp, dp, C.c_uint32, C.c_float, vp, C.POINTER(vp),
                                            C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64), vp]
    L.b200icp_fill_no_key.argtypes = [vp, vp, C.c_size_t]
    L.b200icp_match.argtypes = [vp, vp, vp, dp, C.POINTER(C.c_uint8), up, up, dp, dp, up]
    L.b200icp_align.argtypes = [vp, vp, vp, dp, C.POINTER(Result)]
    L.b200icp_align_with.argtypes = [vp, vp, vp, dp, C.POINTER(CallParams), C.POINTER(Result)]
    L.b200icp_call_params_of.argtypes = [C.POINTER(Params), C.POINTER(CallParams)]
    L.b200icp_call_params_of.restype = None
    L.b200icp_comm_unique_id.argtypes = [C.c_char_p]
    L.b200icp_comm_create.argtypes = [vp, C.c_char_p, C.c_int, C.c_int, C.POINTER(vp)]
    L.b200icp_comm_destroy.argtypes = [vp]
    L.b200icp_comm_destroy.restype = None
    L.b200icp_sharded_map_create.argtypes = [vp, vp, vp, vp, C.c_size_t, C.c_float, C.c_int, C.c_float, C.POINTER(vp)]
    L.b200icp_sharded_map_destroy.argtypes = [vp]
    L.b200icp_sharded_map_destroy.restype = None
    L.b200icp_sharded_map_local_size.argtypes = [vp]
    L.b200icp_sharded_map_local_size.restype = C.c_size_t
    L.b200icp_sharded_knn_keys.argtypes = [vp, vp, dp, C.c_uint32, C.c_float, vp]
    L.b200icp_sharded_align.argtypes = [vp, vp, dp, C.POINTER(CallParams), C.POINTER(Result)]
    L.b200icp_align_batch.argtypes = [vp, C.c_size_t, C.POINTER(vp), C.POINTER(vp), dp,
                                      C.POINTER(Result)]
    L.b200icp_profile_enable.argtypes = [vp, C.c_int]
    L.b200icp_profile_enable.restype = None
    L.b200icp_profile_reset.argtypes = [vp]
    L.b200icp_profile_reset.restype = None
    L.b200icp_profile_get.argtypes = [vp, C.POINTER(Profile)]
    L.b200icp_profile_get.restype = None
    L.b200icp_synchronize.argtypes = [vp]
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise B200IcpError(f"b200icp error {rc}: {lib().b200icp_last_error().decode()}")


def _ptr(a, ct):
    return a.ctypes.data_as(C.POINTER(ct))


def default_params(**kw):
    p = Params()
    lib().b200icp_default_params(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def params_from_yaml(text):
    p = Params()
    _check(lib().b200icp_params_from_yaml(text.encode(), C.byref(p)))
    return p


class EdgesPlanesParams(C.Structure):
    """b200icp_edges_planes_params_t (include/b200icp.h)."""
    _fields_ = [("voxel_filter_resolution", C.c_float), ("full_pointcloud_decimation", C.c_uint32),
                ("voxel_filter_decimation", C.c_uint32), ("voxel_filter_max_e2_e0", C.c_float),
                ("voxel_filter_max_e1_e0", C.c_float), ("voxel_filter_min_e2_e0", C.c_float),
                ("voxel_filter_min_e1_e0", C.c_float), ("min_points_per_voxel", C.c_uint32)]


class Cloud:
    """A point layer resident in HBM with its search index."""

    def __init__(self, icp, handle):
        self.icp = icp
        self.h = handle

    def __len__(self):
        return lib().b200icp_cloud_size(self.h)

    def download(self):
        n = len(self)
        x, y, z = (np.empty(n, dtype=np.float32) for _ in range(3))
        if n:
            _check(lib().b200icp_cloud_download(self.h, _ptr(x, C.c_float), _ptr(y, C.c_float),
                                                _ptr(z, C.c_float)))
        return np.stack([x, y, z], axis=1)

    def free(self):
        if self.h:
            if self.icp is not None and self.icp.h:  # the context must outlive its clouds
                lib().b200icp_cloud_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class ICP:
    """The mp2p_icp::ICP object of one ICP settings block, on one GPU."""

    def __init__(self, params=None, yaml_text=None, device=0):
        L = lib()
        self.h = C.c_void_p()
        if yaml_text is not None:
            _check(L.b200icp_create_from_yaml(yaml_text.encode(), device, C.byref(self.h)))
        else:
            p = params if params is not None else default_params()
            _check(L.b200icp_create(C.byref(p), device, C.byref(self.h)))
        self.device = device

    @property
    def params(self):
        p = Params()
        _check(lib().b200icp_get_params(self.h, C.byref(p)))
        return p

    def close(self):
        if self.h:
            lib().b200icp_destroy(self.h)
            self.h = None

    # ---- clouds
    def upload(self, xyz=None, x=None, y=None, z=None, search_radius=0.0):
        """Host SoA (or an (N,3) array, split here) -> indexed cloud in HBM."""
        if xyz is not None:
            xyz = np.asarray(xyz, dtype=np.float32).reshape(-1, 3)
            x, y, z = (np.ascontiguousarray(xyz[:, i]) for i in range(3))
        n = len(x)
        h = C.c_void_p()
        _check(lib().b200icp_cloud_upload(self.h, x.ctypes.data, y.ctypes.data, z.ctypes.data, n,
                                          search_radius, C.byref(h)))
        return Cloud(self, h)

    def upload_raw(self, xyz):
        """Coordinates only, no search index (b200icp_cloud_upload_raw): the input of voxel_decimate."""
        xyz = np.asarray(xyz, dtype=np.float32).reshape(-1, 3)
        x, y, z = (np.ascontiguousarray(xyz[:, i]) for i in range(3))
        h = C.c_void_p()
        _check(lib().b200icp_cloud_upload_raw(self.h, x.ctypes.data, y.ctypes.data, z.ctypes.data, len(x), C.byref(h)))
        return Cloud(self, h)

    def upload_ptrs(self, px, py, pz, n, search_radius=0.0):
        """Host pointers (e.g. pinned torch tensors' data_ptr())."""
        h = C.c_void_p()
        _check(lib().b200icp_cloud_upload(self.h, px, py, pz, n, search_radius, C.byref(h)))
        return Cloud(self, h)

    def from_device(self, dx, dy, dz, n, search_radius=0.0):
        """Device pointers (inputs already resident in HBM)."""
        h = C.c_void_p()
        _check(lib().b200icp_cloud_from_device(self.h, dx, dy, dz, n, search_radius, C.byref(h)))
        return Cloud(self, h)

    def voxel_decimate(self, cloud, resolution, use_average=False, search_radius=0.0,
                       want_indices=False):
        h = C.c_void_p()
        keep = np.empty(max(len(cloud), 1), dtype=np.uint32) if want_indices else None
        _check(lib().b200icp_voxel_decimate(self.h, cloud.h, resolution, int(use_average),
                                            search_radius, C.byref(h),
                                            _ptr(keep, C.c_uint32) if want_indices else None))
        out = Cloud(self, h)
        if want_indices:
            return out, keep[:len(out)].copy()
        return out

    def filter_edges_planes(self, cloud, search_radius=0.0, **kw):
        """FilterEdgesPlanes: (edges, planes, full_decim) clouds, per-point layer flags (bit 0 / 1 / 2) and the
        number of classified voxels; keyword arguments override the shipped parameter values."""
        prm = EdgesPlanesParams()
        lib().b200icp_edges_planes_defaults(C.byref(prm))
        for k, v in kw.items():
            if not hasattr(prm, k):
                raise TypeError(f"unknown FilterEdgesPlanes parameter {k}")
            setattr(prm, k, v)
        hs = (C.c_void_p * 3)()
        flags = np.zeros(max(len(cloud), 1), dtype=np.uint8)
        nv = C.c_uint32(0)
        _check(lib().b200icp_filter_edges_planes(self.h, cloud.h, C.byref(prm), search_radius, C.byref(hs),
                                                 _ptr(flags, C.c_uint8), C.byref(nv)))
        return [Cloud(self, C.c_void_p(h)) for h in hs], flags[:len(cloud)], int(nv.value)

    # ---- search / matching / registration
    def knn(self, ref, queries, k, max_dist, pose6=None):
        nq = len(queries)
        idx = np.empty((nq, k), dtype=np.uint32)
        d2 = np.empty((nq, k), dtype=np.float32)
        pose = None if pose6 is None else np.ascontiguousarray(pose6, dtype=np.float64)
        _check(lib().b200icp_knn(self.h, ref.h, queries.h,
                                 None if pose is None else _ptr(pose, C.c_double), k, max_dist,
                                 _ptr(idx, C.c_uint32), _ptr(d2, C.c_float)))
        return idx, d2

    def knn_keys_device(self, ref, queries, k, max_dist, d_keys_out, d_index_map=0, pose6=None):
        """Same search, packed keys (d2 bits << 32 | index) left on the device at
        the raw pointer d_keys_out [len(queries) * k] uint64; d_index_map: device
        pointer of the shard-local -> global index table (0: none)."""
        pose = None if pose6 is None else np.ascontiguousarray(pose6, dtype=np.float64)
        _check(lib().b200icp_knn_keys_device(self.h, ref.h, queries.h,
                                             None if pose is None else _ptr(pose, C.c_double), k, max_dist,
                                             d_index_map or None, d_keys_out))

    def knn_keys_scatter(self, ref, queries, k, max_dist, gather_ptrs, rank, d_index_map=0, atomic_min=False,
                         pose6=None):
        """Fused search + exchange: rows (or, k = 1 with atomic_min, the folded
        minimum) stored straight into every rank's gather buffer over NVLink.
        gather_ptrs: device pointers of the ranks' buffers as mapped here."""
        pose = None if pose6 is None else np.ascontiguousarray(pose6, dtype=np.float64)
        arr = (C.c_void_p * len(gather_ptrs))(*gather_ptrs)
        _check(lib().b200icp_knn_keys_scatter(self.h, ref.h, queries.h,
                                              None if pose is None else _ptr(pose, C.c_double), k, max_dist,
                                              d_index_map or None, arr, len(gather_ptrs), rank,
                                              int(bool(atomic_min))))

    def peer_alloc(self, nbytes):
        """(device pointer, 64-byte IPC handle) of an exchange buffer other ranks can map."""
        p = C.c_void_p()
        h = C.create_string_buffer(64)
        _check(lib().b200icp_peer_alloc(self.h, nbytes, C.byref(p), h))
        return p.value, h.raw

    def peer_open(self, handle):
        p = C.c_void_p()
        _check(lib().b200icp_peer_open(self.h, handle, C.byref(p)))
        return p.value

    def peer_close(self, ptr):
        _check(lib().b200icp_peer_close(self.h, ptr))

    def peer_free(self, ptr):
        _check(lib().b200icp_peer_free(self.h, ptr))

    def peer_barrier(self, flag_ptrs, rank, epoch):
        """Barrier of the node's ranks through peer memory (no collective library)."""
        arr = (C.c_void_p * len(flag_ptrs))(*flag_ptrs)
        _check(lib().b200icp_peer_barrier(self.h, arr, len(flag_ptrs), rank, epoch))

    def knn_keys_exchange(self, ref, queries, k, max_dist, base_ptrs, rank, epoch, d_out, d_index_map=0,
                          pose6=None):
        """The whole sharded query in one call (see b200icp_knn_keys_exchange);
        returns the advanced barrier epoch."""
        pose = None if pose6 is None else np.ascontiguousarray(pose6, dtype=np.float64)
        arr = (C.c_void_p * len(base_ptrs))(*base_ptrs)
        ep = C.c_uint64(epoch)
        _check(lib().b200icp_knn_keys_exchange(self.h, ref.h, queries.h,
                                               None if pose is None else _ptr(pose, C.c_double), k, max_dist,
                                               d_index_map or None, arr, len(base_ptrs), rank, C.byref(ep),
                                               d_out))
        return ep.value

    def fill_no_key(self, d_keys, n):
        _check(lib().b200icp_fill_no_key(self.h, d_keys, n))

    def merge_keys_device(self, d_parts, parts, part_stride, nq, k, d_out):
        """k smallest of `parts` ascending key lists per query, device pointers."""
        _check(lib().b200icp_merge_keys_device(self.h, d_parts, parts, part_stride, nq, k, d_out))

    def match(self, from_global, to_local, pose6=None):
        # Matcher_Points_DistanceThreshold pairs with the single nearest neighbour
        n, k = len(to_local), (1 if self.params.matcher_kind == 1 else self.params.knn)
        paired = np.zeros(n, dtype=np.uint8)
        nn_idx = np.empty((n, k), dtype=np.uint32)
        nn_cnt = np.zeros(n, dtype=np.uint32)
        cen = np.zeros((n, 3))
        nor = np.zeros((n, 3))
        npair = C.c_uint32(0)
        pose = None if pose6 is None else np.ascontiguousarray(pose6, dtype=np.float64)
        _check(lib().b200icp_match(self.h, from_global.h, to_local.h,
                                   None if pose is None else _ptr(pose, C.c_double),
                                   _ptr(paired, C.c_uint8), _ptr(nn_idx, C.c_uint32),
                                   _ptr(nn_cnt, C.c_uint32), _ptr(cen, C.c_double),
                                   _ptr(nor, C.c_double), C.byref(npair)))
        return dict(n=npair.value, paired=paired, nn_idx=nn_idx, nn_cnt=nn_cnt, centroid=cen,
                    normal=nor)

    def align(self, from_global, to_local, guess6=None):
        g = np.ascontiguousarray(np.zeros(6) if guess6 is None else guess6, dtype=np.float64)
        r = Result()
        _check(lib().b200icp_align(self.h, from_global.h, to_local.h, _ptr(g, C.c_double),
                                   C.byref(r)))
        return r.as_dict()

    def align_with(self, from_global, to_local, guess6=None, **call):
        """b200icp_align_with: the object's matchers / solvers / quality evaluators with this call's own
        mp2p_icp::Parameters; keyword arguments override fields of the object's parameter block."""
        g = np.ascontiguousarray(np.zeros(6) if guess6 is None else guess6, dtype=np.float64)
        base = Params()
        _check(lib().b200icp_get_params(self.h, C.byref(base)))
        cp = CallParams()
        lib().b200icp_call_params_of(C.byref(base), C.byref(cp))
        for k, v in call.items():
            if not hasattr(cp, k):
                raise KeyError(k)
            setattr(cp, k, v)
        r = Result()
        _check(lib().b200icp_align_with(self.h, from_global.h, to_local.h, _ptr(g, C.c_double), C.byref(cp),
                                        C.byref(r)))
        return r.as_dict()

    def align_batch(self, from_list, to_list, guesses):
        n = len(from_list)
        g = np.ascontiguousarray(guesses, dtype=np.float64).reshape(n, 6)
        fa = (C.c_void_p * n)(*[c.h for c in from_list])
        ta = (C.c_void_p * n)(*[c.h for c in to_list])
        res = (Result * n)()
        _check(lib().b200icp_align_batch(self.h, n, fa, ta, _ptr(g, C.c_double), res))
        return [r.as_dict() for r in res]

    # ---- measurement
    def profile_enable(self, on=True):
        lib().b200icp_profile_enable(self.h, int(on))

    def profile_reset(self):
        lib().b200icp_profile_reset(self.h)

    def profile(self):
        p = Profile()
        lib().b200icp_profile_get(self.h, C.byref(p))
        return p.as_dict()

    @staticmethod
    def profile_of_handle(raw_handle):
        """Profile counters of a raw b200icp_t* (e.g. LidarOdometry.icp_handle())."""
        p = Profile()
        lib().b200icp_profile_get(raw_handle, C.byref(p))
        return p.as_dict()

    def synchronize(self):
        _check(lib().b200icp_synchronize(self.h))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def device_count():
    return lib().b200icp_device_count()


# ---- a map sharded over the GPUs of one box, natively (the library owns the NCCL communicator) --------------
COMM_ID_BYTES = 128


def comm_unique_id():
    """128 bytes for b200icp_comm_create; made by ONE rank and handed to the others by any means."""
    buf = C.create_string_buffer(COMM_ID_BYTES)
    _check(lib().b200icp_comm_unique_id(buf))
    return buf.raw


class Comm:
    def __init__(self, icp, unique_id, world, rank):
        self.icp, self.world, self.rank = icp, world, rank
        h = C.c_void_p()
        _check(lib().b200icp_comm_create(icp.h, unique_id, world, rank, C.byref(h)))
        self.h = h

    def close(self):
        if self.h:
            lib().b200icp_comm_destroy(self.h)
            self.h = None


class NativeShardedMap:
    """b200icp_sharded_map_*: every rank passes the whole map, keeps the cells it owns."""

    def __init__(self, comm, xyz, cell=4.0, interleaved=True, search_radius=0.0):
        self.comm = comm
        xyz = np.asarray(xyz, dtype=np.float32).reshape(-1, 3)
        x, y, z = (np.ascontiguousarray(xyz[:, i]) for i in range(3))
        h = C.c_void_p()
        _check(lib().b200icp_sharded_map_create(comm.h, x.ctypes.data, y.ctypes.data, z.ctypes.data, len(x), cell,
                                                int(interleaved), search_radius, C.byref(h)))
        self.h = h

    def local_size(self):
        return int(lib().b200icp_sharded_map_local_size(self.h))

    def knn_keys(self, queries, k, max_dist, d_keys_ptr, pose6=None):
        p = None if pose6 is None else _ptr(np.ascontiguousarray(pose6, dtype=np.float64), C.c_double)
        _check(lib().b200icp_sharded_knn_keys(self.h, queries.h, p, k, max_dist, d_keys_ptr))

    def align(self, to_local, guess6=None):
        g = np.ascontiguousarray(np.zeros(6) if guess6 is None else guess6, dtype=np.float64)
        r = Result()
        _check(lib().b200icp_sharded_align(self.h, to_local.h, _ptr(g, C.c_double), None, C.byref(r)))
        return r.as_dict()

    def close(self):
        if self.h:
            lib().b200icp_sharded_map_destroy(self.h)
            self.h = None
