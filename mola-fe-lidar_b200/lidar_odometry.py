"""ctypes binding of lib/libmola_fe_lidar_b200.so -- the host-side mirror of
mola::LidarOdometry (include/b200_lidar_odometry.h): initialize from the
reference's YAML parameter files, feed observations, read the state back.
No CPU fallback: creation fails when the CUDA library cannot reach a B200."""
import ctypes as C
import os
import time

import numpy as np

from . import capi

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "lib", "libmola_fe_lidar_b200.so")
PARAMS_DIR = os.path.join(PKG_DIR, "params")

EXPORTS = [
    "b200lo_last_error", "b200lo_create", "b200lo_destroy", "b200lo_reset",
    "b200lo_on_new_observation", "b200lo_enqueue_observation", "b200lo_queue_length",
    "b200lo_process_observation", "b200lo_spin_once",
    "b200lo_wait_idle", "b200lo_get_state", "b200lo_get_factors", "b200lo_dump_params",
    "b200lo_dump_profile", "b200lo_icp_handle", "b200lo_last_montecarlo",
]


class State(C.Structure):
    _fields_ = [
        ("last_obs_tim", C.c_double),
        ("accum_since_last_kf", C.c_double * 6),
        ("last_twist", C.c_double * 6),
        ("last_iter_twist_is_good", C.c_int32),
        ("last_kf", C.c_uint64),
        ("n_keyframes", C.c_uint64), ("n_factors", C.c_uint64), ("n_localizations", C.c_uint64),
        ("n_processed", C.c_uint64), ("n_dropped", C.c_uint64), ("n_icp", C.c_uint64),
        ("last_icp_goodness", C.c_double),
        ("last_icp_pose", C.c_double * 6),
        ("last_icp_iterations", C.c_uint32), ("last_icp_termination", C.c_uint32),
        ("last_points_size", C.c_size_t),
        ("n_graph_edges", C.c_uint64), ("n_checked_pairs", C.c_uint64),
        ("n_kf_spills", C.c_uint64), ("n_kf_reloads", C.c_uint64),
    ]


class Factor(C.Structure):
    _fields_ = [("from_kf", C.c_uint64), ("to_kf", C.c_uint64), ("rel_pose", C.c_double * 6)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise capi.B200IcpError(f"{LIB_PATH} not found: run __graft_entry__.build() (no CPU fallback)")
    capi.lib()  # dependency, loaded first from the same directory
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.b200lo_last_error.restype = C.c_char_p
    L.b200lo_create.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(vp)]
    L.b200lo_destroy.argtypes = [vp]
    L.b200lo_destroy.restype = None
    L.b200lo_reset.argtypes = [vp]
    L.b200lo_reset.restype = None
    for name in ("b200lo_on_new_observation", "b200lo_process_observation", "b200lo_enqueue_observation"):
        getattr(L, name).argtypes = [vp, C.c_char_p, C.c_double, vp, vp, vp, C.c_size_t]
    L.b200lo_queue_length.argtypes = [vp]
    L.b200lo_queue_length.restype = C.c_size_t
    L.b200lo_spin_once.argtypes = [vp]
    L.b200lo_spin_once.restype = None
    L.b200lo_wait_idle.argtypes = [vp]
    L.b200lo_wait_idle.restype = None
    L.b200lo_get_state.argtypes = [vp, C.POINTER(State)]
    L.b200lo_get_factors.argtypes = [vp, C.POINTER(Factor), C.c_size_t]
    L.b200lo_get_factors.restype = C.c_size_t
    L.b200lo_dump_params.argtypes = [vp, C.c_char_p, C.c_size_t]
    L.b200lo_dump_params.restype = C.c_size_t
    L.b200lo_dump_profile.argtypes = [vp, C.c_char_p, C.c_size_t]
    L.b200lo_dump_profile.restype = C.c_size_t
    L.b200lo_last_montecarlo.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), vp, vp, C.c_size_t,
                                         C.POINTER(C.c_double), vp]
    L.b200lo_last_montecarlo.restype = C.c_size_t
    L.b200lo_icp_handle.argtypes = [vp, C.c_int]
    L.b200lo_icp_handle.restype = vp
    _lib = L
    return L


def system_yaml(params_file="kitti-default.yaml", sensor_label="lidar", voxel_resolution=None,
                voxel_average=False, extra=""):
    """The module block a MOLA SLAM-system file would hold for this front-end:
    `params:` pulls the reference-format parameter file in with $include{};
    the `pointcloud_filter` block the shipped YAML leaves undefined is added
    here when a voxel resolution is given."""
    path = params_file if os.path.isabs(params_file) else os.path.join(PARAMS_DIR, params_file)
    txt = f"raw_sensor_label: {sensor_label}\nparams:\n  $include{{{path}}}\n"
    if voxel_resolution:
        txt += ("  pointcloud_filter:\n"
                "    - class_name: mp2p_icp_filters::FilterDecimateVoxels\n"
                "      params:\n"
                f"        voxel_filter_resolution: {voxel_resolution}\n"
                f"        use_voxel_average: {'true' if voxel_average else 'false'}\n")
    return txt + extra


class LidarOdometry:
    """mola::LidarOdometry: initialize(yaml) / onNewObservation / spinOnce / reset."""

    def __init__(self, yaml_text=None, yaml_path=None, mola_dir=None):
        L = lib()
        self.h = C.c_void_p()
        if yaml_text is None and yaml_path is None:
            yaml_text = system_yaml()
        rc = L.b200lo_create(yaml_path.encode() if yaml_path else None,
                             yaml_text.encode() if yaml_text else None,
                             mola_dir.encode() if mola_dir else None, C.byref(self.h))
        if rc != 0:
            raise capi.B200IcpError(f"LidarOdometry.initialize failed ({rc}): {L.b200lo_last_error().decode()}")

    def onNewObservation(self, xyz, timestamp, label="lidar", sync=False):
        xyz = np.asarray(xyz, dtype=np.float32).reshape(-1, 3)
        x, y, z = (np.ascontiguousarray(xyz[:, i]) for i in range(3))
        return self.onNewObservationSoA(x.ctypes.data, y.ctypes.data, z.ctypes.data, len(x), timestamp,
                                        label, sync, _keep=(x, y, z))

    def onNewObservationSoA(self, px, py, pz, n, timestamp, label="lidar", sync=True, _keep=None):
        f = lib().b200lo_process_observation if sync else lib().b200lo_on_new_observation
        rc = f(self.h, label.encode(), timestamp, px, py, pz, n)
        if rc != 0:
            raise capi.B200IcpError(f"onNewObservation failed ({rc}): {lib().b200lo_last_error().decode()}")

    def enqueueObservationSoA(self, px, py, pz, n, timestamp, label="lidar", max_queued=4):
        """Asynchronous feed without a copy (b200lo_enqueue_observation): the buffers must stay valid until the
        scan has been processed.  Waits while more than `max_queued` scans are pending, so that a harness feeding
        faster than real time never triggers the >10-queued drop rule (LidarOdometry.cpp:171-179)."""
        L = lib()
        while L.b200lo_queue_length(self.h) > max_queued:
            time.sleep(20e-6)
        rc = L.b200lo_enqueue_observation(self.h, label.encode(), timestamp, px, py, pz, n)
        if rc != 0:
            raise capi.B200IcpError(f"enqueueObservation failed ({rc}): {L.b200lo_last_error().decode()}")

    def spinOnce(self):
        lib().b200lo_spin_once(self.h)

    def reset(self):
        lib().b200lo_reset(self.h)

    def wait_idle(self):
        lib().b200lo_wait_idle(self.h)

    def state(self):
        s = State()
        lib().b200lo_get_state(self.h, C.byref(s))
        d = {}
        for k, _ in State._fields_:
            v = getattr(s, k)
            d[k] = np.array(v) if hasattr(v, "__len__") else v
        return d

    def factors(self):
        n = lib().b200lo_get_factors(self.h, None, 0)
        arr = (Factor * max(n, 1))()
        n = lib().b200lo_get_factors(self.h, arr, n)
        return [(f.from_kf, f.to_kf, np.array(f.rel_pose)) for f in arr[:n]]

    def last_montecarlo(self):
        """The last loop-closure attempt: dict(from_kf, to_kf, guesses [n,6], goodness [n], best_goodness,
        best_pose) or None."""
        cap = 64
        g, q, bp = np.zeros((cap, 6)), np.zeros(cap), np.zeros(6)
        a, b, bg = C.c_uint64(), C.c_uint64(), C.c_double()
        n = lib().b200lo_last_montecarlo(self.h, C.byref(a), C.byref(b), g.ctypes.data, q.ctypes.data, cap, C.byref(bg),
                                         bp.ctypes.data)
        if n == 0:
            return None
        return dict(from_kf=int(a.value), to_kf=int(b.value), guesses=g[:n].copy(), goodness=q[:n].copy(),
                    best_goodness=float(bg.value), best_pose=bp)

    def _dump(self, fn):
        n = fn(self.h, None, 0)
        buf = C.create_string_buffer(n + 1)
        fn(self.h, buf, n + 1)
        return buf.value.decode()

    def params(self):
        out = {}
        for line in self._dump(lib().b200lo_dump_params).splitlines():
            k, _, v = line.partition("=")
            out[k] = v
        return out

    def profile(self):
        out = {}
        for line in self._dump(lib().b200lo_dump_profile).splitlines():
            name, cnt, tot, mx = line.rsplit(",", 3)
            out[name] = (int(cnt), float(tot), float(mx))
        return out

    def icp_handle(self, kind=0):
        return lib().b200lo_icp_handle(self.h, kind)

    def close(self):
        if self.h:
            lib().b200lo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
