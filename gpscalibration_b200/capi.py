"""ctypes binding of the C ABI in include/loamgpu.h (libloamgpu.so, built in-tree under csrc/).

This is the boundary a maintainer of the reference binds to; nothing here computes anything.  If the shared library
is missing the import of the package still works but every call raises (no CPU fallback).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

LOAM_OK, LOAM_EINVAL, LOAM_ECUDA, LOAM_ENOSPC, LOAM_ESTATE, LOAM_EUNSUPPORTED = 0, -1, -2, -3, -4, -5

CLOUD = dict(full=0, sharp=1, less_sharp=2, flat=3, less_flat=4, corner_last=5, surf_last=6, full_res3=7, corner_stack=8,
             surf_stack=9, corner_map=10, surf_map=11, surround=12, registered=13)
DIAG = dict(curvature=(0, np.float32), picked_mask=(1, np.uint8), label=(2, np.int8), scan_start=(3, np.int32),
            scan_end=(4, np.int32))


class LoamError(RuntimeError):
    def __init__(self, code, where, detail=""):
        self.code = code
        super().__init__(f"{where}: error {code} ({detail})")


class Params(C.Structure):
    _fields_ = [("n_scans", C.c_int), ("ring_mode", C.c_int), ("ring_ang_min", C.c_float), ("ring_ang_step", C.c_float),
                ("skip_frame_num", C.c_int), ("max_points", C.c_int), ("max_map_points", C.c_int),
                ("want_registered", C.c_int), ("want_surround", C.c_int), ("pose_message_hop", C.c_int), ("gn_max_ctas", C.c_int)]


class Counts(C.Structure):
    _fields_ = [("n_full", C.c_int), ("n_sharp", C.c_int), ("n_less_sharp", C.c_int), ("n_flat", C.c_int),
                ("n_less_flat", C.c_int)]


class OdomResult(C.Structure):
    _fields_ = [("transform_sum", C.c_float * 6), ("transformation", C.c_float * 6), ("odom_published", C.c_int),
                ("clouds_published", C.c_int), ("fullres_published", C.c_int), ("iterations", C.c_int),
                ("n_corner_last", C.c_int), ("n_surf_last", C.c_int)]


class MapResult(C.Structure):
    _fields_ = [("transform_aft_mapped", C.c_float * 6), ("transform_bef_mapped", C.c_float * 6),
                ("transform_tobe_mapped", C.c_float * 6), ("optimised", C.c_int), ("iterations", C.c_int),
                ("surround_published", C.c_int), ("n_corner_stack", C.c_int), ("n_surf_stack", C.c_int),
                ("n_corner_map", C.c_int), ("n_surf_map", C.c_int), ("n_surround", C.c_int), ("n_registered", C.c_int)]


class SweepResult(C.Structure):
    _fields_ = [("counts", Counts), ("odom", OdomResult), ("map", MapResult), ("mapping_ran", C.c_int)]


# every symbol include/loamgpu.h declares (tests check the library exports all of them)
SYMBOLS = ["loam_strerror", "loam_last_cuda_error", "loam_default_params", "loam_create", "loam_destroy", "loam_reset",
           "loam_stream", "loam_launch_count", "loam_stats", "loam_profile", "loam_profile_read", "loam_host_times", "loam_launch_latency", "loam_pose_message_hop", "loam_imu_push", "loam_get_imu_trans", "loam_extract", "loam_extract_device", "loam_extract_batch", "loam_odometry_process", "loam_odometry_process_batch",
           "loam_mapping_odometry", "loam_mapping_process", "loam_integrate_odometry", "loam_integrate_mapping", "loam_process_sweep", "loam_process_sweep_device",
           "loam_get_cloud", "loam_get_cloud_wire", "loam_get_diag", "loam_voxel_grid", "loam_odom_set_inputs", "loam_odom_iter",
           "loam_odom_get_corr", "loam_transform_to_end", "loam_map_set_inputs", "loam_map_iter", "loam_map_get_corr",
           "loam_gn_solve", "loam_map_iter_partial", "loam_map_finish_reduced", "loam_shard_export", "loam_shard_connect", "loam_shard_set_slab", "loam_shard_inject",
           "loam_map_iter_allreduce", "loam_map_optimize", "loam_pipeline_create", "loam_pipeline_destroy",
           "loam_pipeline_reset", "loam_pipeline_last_error", "loam_pipeline_submit", "loam_pipeline_submit_batch", "loam_pipeline_submit_lockstep", "loam_pipeline_submit_device", "loam_pipeline_imu_push", "loam_pipeline_wait", "loam_pipeline_pending",
           "loam_pipeline_stream",
           "loam_pipeline_stats", "loam_pipeline_stage_times", "loam_pipeline_handle", "loam_replay_segments", "loam_track_svd3", "loam_track_speed_weights", "loam_track_residual_weights",
           "loam_track_icp", "loam_track_smooth", "loam_track_calibrate", "loam_track_calibrate_long"]


def pose_message_hop(transform_sum):
    """loam_pose_message_hop: the odometry pose as it arrives after the quaternion message (LO:1066-1078 -> LM:322-332)."""
    a = _f32(transform_sum)
    out = np.zeros(6, np.float32)
    rc = load_library().loam_pose_message_hop(a.ctypes.data, out.ctypes.data)
    if rc:
        raise LoamError(rc, "loam_pose_message_hop")
    return out


def library_path():
    return os.path.join(_HERE, "csrc", "libloamgpu.so")


def load_library():
    """Loads libloamgpu.so (fails loudly when it has not been built)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise RuntimeError(f"{path} missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
    lib = C.CDLL(path)
    vp, ip, fp, dp = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_double)
    lib.loam_strerror.restype = C.c_char_p
    lib.loam_strerror.argtypes = [C.c_int]
    lib.loam_last_cuda_error.restype = C.c_char_p
    lib.loam_last_cuda_error.argtypes = [vp]
    lib.loam_default_params.argtypes = [C.POINTER(Params)]
    lib.loam_create.argtypes = [C.POINTER(Params), C.c_int, C.POINTER(vp)]
    lib.loam_destroy.argtypes = [vp]
    lib.loam_reset.argtypes = [vp]
    lib.loam_stream.restype = vp
    lib.loam_stream.argtypes = [vp]
    lib.loam_launch_count.restype = C.c_longlong
    lib.loam_launch_count.argtypes = [vp]
    lib.loam_stats.argtypes = [vp, vp]
    lib.loam_profile.argtypes = [vp, C.c_int]
    lib.loam_host_times.argtypes = [vp, vp, C.c_int]
    lib.loam_pose_message_hop.argtypes = [vp, vp]
    lib.loam_launch_latency.argtypes = [vp, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.loam_profile_read.argtypes = [vp, vp, vp, vp, C.c_int]
    lib.loam_imu_push.argtypes = [vp, C.c_double, vp, vp, vp]
    lib.loam_get_imu_trans.argtypes = [vp, vp]
    lib.loam_extract.argtypes = [vp, vp, C.c_int, C.c_int, C.c_double, vp, C.POINTER(Counts)]
    lib.loam_extract_device.argtypes = [vp, vp, C.c_int, C.c_int, C.c_double, vp, C.POINTER(Counts)]
    lib.loam_extract_batch.argtypes = [vp, C.c_int, vp, vp, C.c_int, vp, vp]
    lib.loam_odometry_process.argtypes = [vp, C.POINTER(OdomResult)]
    lib.loam_odometry_process_batch.argtypes = [vp, C.c_int, vp]
    lib.loam_mapping_odometry.argtypes = [vp, vp]
    lib.loam_mapping_process.argtypes = [vp, C.POINTER(MapResult)]
    lib.loam_integrate_odometry.argtypes = [vp, vp, C.c_double, vp, vp]
    lib.loam_integrate_mapping.argtypes = [vp, vp, vp]
    lib.loam_process_sweep.argtypes = [vp, vp, C.c_int, C.c_int, C.c_double, C.POINTER(SweepResult)]
    lib.loam_process_sweep_device.argtypes = [vp, vp, C.c_int, C.c_int, C.c_double, C.POINTER(SweepResult)]
    lib.loam_get_cloud.argtypes = [vp, C.c_int, vp, C.c_int, ip]
    lib.loam_get_cloud_wire.argtypes = [vp, C.c_int, vp, C.c_int, ip]
    lib.loam_get_diag.argtypes = [vp, C.c_int, vp, C.c_int, ip]
    lib.loam_voxel_grid.argtypes = [vp, vp, C.c_int, C.c_float, vp, C.c_int, ip]
    lib.loam_odom_set_inputs.argtypes = [vp, vp, C.c_int, vp, C.c_int, vp, C.c_int, vp, C.c_int]
    lib.loam_odom_iter.argtypes = [vp, C.c_int, vp, vp, vp, ip]
    lib.loam_odom_get_corr.argtypes = [vp, vp, vp, C.c_int, vp, vp, vp, C.c_int]
    lib.loam_transform_to_end.argtypes = [vp, vp, C.c_int, vp, vp, vp]
    lib.loam_map_set_inputs.argtypes = [vp, vp, C.c_int, vp, C.c_int, vp, C.c_int, vp, C.c_int]
    lib.loam_map_iter.argtypes = [vp, C.c_int, vp, vp, vp, ip]
    lib.loam_map_get_corr.argtypes = [vp, vp, C.c_int, vp, C.c_int]
    lib.loam_gn_solve.argtypes = [vp, vp, C.c_int, C.c_float, vp, vp]
    lib.loam_map_iter_partial.argtypes = [vp, C.c_int, vp, vp]
    lib.loam_map_finish_reduced.argtypes = [vp, vp, vp, ip]
    lib.loam_shard_export.argtypes = [vp, vp]
    lib.loam_shard_connect.argtypes = [vp, vp, C.c_int, C.c_int]
    lib.loam_map_iter_allreduce.argtypes = [vp, C.c_int, vp, vp, vp, ip]
    lib.loam_map_optimize.argtypes = [vp, vp, C.c_int, ip]
    lib.loam_shard_set_slab.argtypes = [vp, C.c_float, C.c_float]
    lib.loam_shard_inject.argtypes = [vp, C.c_int, vp]
    lib.loam_pipeline_create.argtypes = [C.POINTER(Params), C.c_int, C.POINTER(vp)]
    lib.loam_pipeline_destroy.argtypes = [vp]
    lib.loam_pipeline_reset.argtypes = [vp]
    lib.loam_pipeline_last_error.restype = C.c_char_p
    lib.loam_pipeline_last_error.argtypes = [vp]
    lib.loam_pipeline_submit.argtypes = [vp, vp, C.c_int, C.c_int, C.c_double]
    lib.loam_pipeline_submit_device.argtypes = [vp, vp, C.c_int, C.c_int, C.c_double]
    lib.loam_pipeline_imu_push.argtypes = [vp, C.c_double, vp, vp, vp]
    lib.loam_pipeline_submit_batch.argtypes = [vp, C.c_int, vp, vp, C.c_int, vp]
    lib.loam_pipeline_submit_lockstep.argtypes = [vp, C.c_int, vp, vp, C.c_int, vp]
    lib.loam_pipeline_wait.argtypes = [vp, C.POINTER(SweepResult)]
    lib.loam_pipeline_pending.argtypes = [vp]
    lib.loam_pipeline_stream.restype = vp
    lib.loam_pipeline_stream.argtypes = [vp, C.c_int]
    lib.loam_pipeline_stats.argtypes = [vp, vp]
    lib.loam_pipeline_stage_times.argtypes = [vp, vp, C.c_int]
    lib.loam_pipeline_handle.restype = vp
    lib.loam_pipeline_handle.argtypes = [vp, C.c_int]
    lib.loam_track_svd3.argtypes = [vp, vp, vp, vp]
    lib.loam_track_speed_weights.argtypes = [vp, C.c_int, vp]
    lib.loam_track_residual_weights.argtypes = [vp, vp, vp, C.c_int, vp]
    lib.loam_track_icp.argtypes = [vp, vp, vp, C.c_int, vp, vp]
    lib.loam_track_smooth.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, vp]
    lib.loam_track_calibrate.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int, vp, vp]
    lib.loam_track_calibrate_long.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]
    _LIB = lib
    return lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class LoamGpu:
    """One handle = one GPU + one stream + the state of the three LOAM stages (thin wrapper, no logic)."""

    def __init__(self, device=0, n_scans=16, ring_mode=0, ring_ang_min=-15.0, ring_ang_step=2.0, skip_frame_num=1,
                 want_registered=False, want_surround=False, max_points=None, max_map_points=None, pose_message_hop=False, gn_max_ctas=0):
        self.lib = load_library()
        p = Params()
        self.lib.loam_default_params(C.byref(p))
        p.n_scans, p.ring_mode, p.ring_ang_min, p.ring_ang_step = n_scans, ring_mode, ring_ang_min, ring_ang_step
        p.skip_frame_num = skip_frame_num
        p.want_registered, p.want_surround = int(want_registered), int(want_surround)
        p.pose_message_hop = int(pose_message_hop)
        p.gn_max_ctas = int(gn_max_ctas)
        if max_points:
            p.max_points = int(max_points)
        if max_map_points:
            p.max_map_points = int(max_map_points)
        self.params = p
        self._h = C.c_void_p()
        self._check(self.lib.loam_create(C.byref(p), device, C.byref(self._h)), "loam_create")

    def _check(self, rc, where):
        if rc != 0:
            detail = self.lib.loam_strerror(rc).decode()
            if rc == LOAM_ECUDA:
                detail += ": " + self.lib.loam_last_cuda_error(self._h).decode()
            raise LoamError(rc, where, detail)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.lib.loam_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- lifecycle
    def reset(self):
        self._check(self.lib.loam_reset(self._h), "loam_reset")

    @property
    def stream(self):
        return self.lib.loam_stream(self._h)

    @property
    def launches(self):
        return int(self.lib.loam_launch_count(self._h))

    def stats(self):
        out = (C.c_longlong * 4)()
        self._check(self.lib.loam_stats(self._h, out), "loam_stats")
        return dict(launches=out[0], h2d_bytes=out[1], d2h_bytes=out[2], syncs=out[3])

    PROFILE_CLASSES = ("extract", "odom_knn", "odom_iter", "to_end", "map_stack", "voxel", "gather", "grid", "map_knn",
                       "map_fit", "insert", "sr_select")

    # host wall-clock sections of the blocking calls (loam_host_times); the names t9..t15 are kept as aliases in tools/
    HOST_SECTIONS = ("extract", "odom_iters", "odom_end", "map_prep", "map_grid", "map_iters", "map_insert", "map_cube_ds", "map_rest",
                     "t9", "t10", "t11", "t12", "t13", "t14", "t15")
    HOST_SECTION_MEANING = {"t9": "cube bookkeeping / grid roll", "t10": "gather of the local map", "t11": "stack voxel grid incl. count read-back",
                            "t12": "arena reserve", "t13": "cube tables + uploads", "t14": "cube merge launches", "t15": "cube merge wait + read-back"}

    def host_times(self, clear=True):
        out = np.zeros(16)
        self._check(self.lib.loam_host_times(self._h, out.ctypes.data, int(clear)), "loam_host_times")
        return {k: float(out[i]) for i, k in enumerate(self.HOST_SECTIONS)}

    def launch_latency(self, n=2000):
        """(us per back-to-back empty launch, us per launch + host-visible completion) on the handle's stream."""
        a, b = C.c_double(), C.c_double()
        self._check(self.lib.loam_launch_latency(self._h, int(n), C.byref(a), C.byref(b)), "loam_launch_latency")
        return a.value, b.value

    def profile(self, enable):
        self._check(self.lib.loam_profile(self._h, int(enable)), "loam_profile")

    def profile_read(self):
        n = len(self.PROFILE_CLASSES)
        ms, units, scopes = np.zeros(n), np.zeros(n), np.zeros(n, np.int64)
        self._check(self.lib.loam_profile_read(self._h, ms.ctypes.data, units.ctypes.data, scopes.ctypes.data, n), "loam_profile_read")
        return {k: dict(ms=float(ms[i]), units=float(units[i]), scopes=int(scopes[i])) for i, k in enumerate(self.PROFILE_CLASSES)}

    # ---- node level
    def extract(self, xyz, stamp=0.0):
        xyz = _f32(xyz)
        c = Counts()
        self._check(self.lib.loam_extract(self._h, xyz.ctypes.data, xyz.shape[0], xyz.strides[0] if xyz.shape[0] > 1 else 12, stamp, None, C.byref(c)),
                    "loam_extract")
        return c

    def imu_push(self, stamp, quat_xyzw, angular_velocity, linear_acceleration):
        """loam_imu_push: one /imu/data message (imuHandler SR:754-837)."""
        q, a, l = (np.ascontiguousarray(v, np.float64) for v in (quat_xyzw, angular_velocity, linear_acceleration))
        self._check(self.lib.loam_imu_push(self._h, float(stamp), q.ctypes.data, a.ctypes.data, l.ctypes.data), "loam_imu_push")

    def imu_trans(self):
        """The 12 floats of /imu_trans of the last extracted sweep (SR:730-745)."""
        out = np.zeros(12, np.float32)
        self._check(self.lib.loam_get_imu_trans(self._h, out.ctypes.data), "loam_get_imu_trans")
        return out

    def extract_device(self, dev_ptr, n, stride_bytes=12, stamp=0.0):
        c = Counts()
        self._check(self.lib.loam_extract_device(self._h, dev_ptr, n, stride_bytes, stamp, None, C.byref(c)), "loam_extract_device")
        return c

    def odometry_process(self):
        r = OdomResult()
        self._check(self.lib.loam_odometry_process(self._h, C.byref(r)), "loam_odometry_process")
        return r

    def mapping_odometry(self, transform_sum):
        t = _f32(transform_sum)
        self._check(self.lib.loam_mapping_odometry(self._h, t.ctypes.data), "loam_mapping_odometry")

    def mapping_process(self):
        r = MapResult()
        self._check(self.lib.loam_mapping_process(self._h, C.byref(r)), "loam_mapping_process")
        return r

    def integrate_odometry(self, transform_sum, stamp):
        t = _f32(transform_sum)
        out, track = np.zeros(6, np.float32), np.zeros(4, np.float64)
        self._check(self.lib.loam_integrate_odometry(self._h, t.ctypes.data, float(stamp), out.ctypes.data, track.ctypes.data),
                    "loam_integrate_odometry")
        return out, track

    def integrate_mapping(self, aft, bef):
        a, b = _f32(aft), _f32(bef)
        self._check(self.lib.loam_integrate_mapping(self._h, a.ctypes.data, b.ctypes.data), "loam_integrate_mapping")

    def process_sweep(self, xyz, stamp=0.0):
        xyz = _f32(xyz)
        r = SweepResult()
        self._check(self.lib.loam_process_sweep(self._h, xyz.ctypes.data, xyz.shape[0], xyz.strides[0] if xyz.shape[0] > 1 else 12, stamp, C.byref(r)),
                    "loam_process_sweep")
        return r

    def process_sweep_device(self, dev_ptr, n, stride_bytes=12, stamp=0.0):
        r = SweepResult()
        self._check(self.lib.loam_process_sweep_device(self._h, dev_ptr, n, stride_bytes, stamp, C.byref(r)),
                    "loam_process_sweep_device")
        return r

    # ---- data access
    def cloud(self, which):
        n = C.c_int()
        w = CLOUD[which] if isinstance(which, str) else which
        self._check(self.lib.loam_get_cloud(self._h, w, None, 0, C.byref(n)), "loam_get_cloud")
        out = np.empty((n.value, 4), np.float32)
        if n.value:
            self._check(self.lib.loam_get_cloud(self._h, w, out.ctypes.data, n.value, C.byref(n)), "loam_get_cloud")
        return out

    # field table of the payload returned by cloud_wire (what pcl::toROSMsg emits for PointXYZI)
    WIRE_POINT_STEP = 32
    WIRE_FIELDS = (("x", 0), ("y", 4), ("z", 8), ("intensity", 16))

    def cloud_wire(self, which):
        """PointCloud2 payload of cloud `which`: uint8 array (n, 32)."""
        n = C.c_int()
        w = CLOUD[which] if isinstance(which, str) else which
        self._check(self.lib.loam_get_cloud_wire(self._h, w, None, 0, C.byref(n)), "loam_get_cloud_wire")
        out = np.empty((n.value, 32), np.uint8)
        if n.value:
            self._check(self.lib.loam_get_cloud_wire(self._h, w, out.ctypes.data, n.value, C.byref(n)), "loam_get_cloud_wire")
        return out

    def extract_wire(self, payload, point_step, x_offset=0, stamp=0.0):
        """loam_extract straight from a PointCloud2 payload (bytes / uint8 array), x y z contiguous at x_offset."""
        buf = np.ascontiguousarray(np.frombuffer(payload, np.uint8) if not isinstance(payload, np.ndarray) else payload.reshape(-1).view(np.uint8))
        n = buf.size // point_step
        c = Counts()
        self._check(self.lib.loam_extract(self._h, buf.ctypes.data + x_offset, n, point_step, stamp, None, C.byref(c)), "loam_extract")
        return c

    # ---- sharded map with the all-reduce fused into the reduction kernel (NVLink peer memory, CUDA IPC)
    def shard_export(self):
        buf = (C.c_ubyte * 64)()
        self._check(self.lib.loam_shard_export(self._h, buf), "loam_shard_export")
        return bytes(buf)

    def shard_connect(self, handles, rank):
        """handles: list of the 64-byte blobs of all ranks in rank order."""
        blob = b"".join(handles)
        arr = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
        self._check(self.lib.loam_shard_connect(self._h, arr, len(handles), rank), "loam_shard_connect")

    def shard_set_slab(self, x_lo, x_hi):
        self._check(self.lib.loam_shard_set_slab(self._h, float(x_lo), float(x_hi)), "loam_shard_set_slab")

    def shard_inject(self, from_rank, sums28):
        s = np.ascontiguousarray(sums28, np.float64)
        self._check(self.lib.loam_shard_inject(self._h, int(from_rank), s.ctypes.data), "loam_shard_inject")

    def map_optimize(self, T, max_iters=10):
        """Whole Gauss-Newton loop on the device (LM:753-1017); returns (T_new, iterations)."""
        T = _f32(T).copy()
        n = C.c_int()
        self._check(self.lib.loam_map_optimize(self._h, T.ctypes.data, int(max_iters), C.byref(n)), "loam_map_optimize")
        return T, n.value

    def map_iter_allreduce(self, it, T):
        T = _f32(T)
        AtA, AtB, n = np.zeros((6, 6), np.float32), np.zeros(6, np.float32), C.c_int()
        self._check(self.lib.loam_map_iter_allreduce(self._h, it, T.ctypes.data, AtA.ctypes.data, AtB.ctypes.data, C.byref(n)),
                    "loam_map_iter_allreduce")
        return AtA, AtB, n.value

    def diag(self, which):
        w, dt = DIAG[which]
        n = C.c_int()
        self._check(self.lib.loam_get_diag(self._h, w, None, 0, C.byref(n)), "loam_get_diag")
        out = np.empty(n.value, dt)
        if n.value:
            self._check(self.lib.loam_get_diag(self._h, w, out.ctypes.data, out.nbytes, C.byref(n)), "loam_get_diag")
        return out

    # ---- stage level
    def voxel_grid(self, pts4, leaf):
        pts4 = _f32(pts4)
        out = np.empty_like(pts4)
        v = C.c_int()
        self._check(self.lib.loam_voxel_grid(self._h, pts4.ctypes.data, pts4.shape[0], leaf, out.ctypes.data, out.shape[0], C.byref(v)),
                    "loam_voxel_grid")
        return out[:v.value].copy()

    def odom_set_inputs(self, sharp, flat, corner_last, surf_last):
        a, b, c, d = _f32(sharp), _f32(flat), _f32(corner_last), _f32(surf_last)
        self._n_sharp, self._n_flat = a.shape[0], b.shape[0]
        self._check(self.lib.loam_odom_set_inputs(self._h, a.ctypes.data, a.shape[0], b.ctypes.data, b.shape[0], c.ctypes.data,
                                                  c.shape[0], d.ctypes.data, d.shape[0]), "loam_odom_set_inputs")

    def odom_iter(self, it, T):
        T = _f32(T)
        AtA, AtB, n = np.zeros((6, 6), np.float32), np.zeros(6, np.float32), C.c_int()
        self._check(self.lib.loam_odom_iter(self._h, it, T.ctypes.data, AtA.ctypes.data, AtB.ctypes.data, C.byref(n)), "loam_odom_iter")
        return AtA, AtB, n.value

    def odom_corr(self, n_sharp, n_flat):
        c1, c2 = np.empty(n_sharp, np.int32), np.empty(n_sharp, np.int32)
        s1, s2, s3 = np.empty(n_flat, np.int32), np.empty(n_flat, np.int32), np.empty(n_flat, np.int32)
        self._check(self.lib.loam_odom_get_corr(self._h, c1.ctypes.data, c2.ctypes.data, n_sharp, s1.ctypes.data, s2.ctypes.data,
                                                s3.ctypes.data, n_flat), "loam_odom_get_corr")
        return c1, c2, s1, s2, s3

    def transform_to_end(self, pts4, T, imu_trans=None):
        pts4, T = _f32(pts4), _f32(T)
        out = np.empty_like(pts4)
        imu = _f32(imu_trans).ctypes.data if imu_trans is not None else None
        self._check(self.lib.loam_transform_to_end(self._h, pts4.ctypes.data, pts4.shape[0], T.ctypes.data, imu, out.ctypes.data),
                    "loam_transform_to_end")
        return out

    def map_set_inputs(self, corner_stack, surf_stack, corner_map, surf_map):
        a, b, c, d = _f32(corner_stack), _f32(surf_stack), _f32(corner_map), _f32(surf_map)
        self._n_cs, self._n_ss = a.shape[0], b.shape[0]
        self._check(self.lib.loam_map_set_inputs(self._h, a.ctypes.data, a.shape[0], b.ctypes.data, b.shape[0], c.ctypes.data,
                                                 c.shape[0], d.ctypes.data, d.shape[0]), "loam_map_set_inputs")

    def map_iter(self, it, T):
        T = _f32(T)
        AtA, AtB, n = np.zeros((6, 6), np.float32), np.zeros(6, np.float32), C.c_int()
        self._check(self.lib.loam_map_iter(self._h, it, T.ctypes.data, AtA.ctypes.data, AtB.ctypes.data, C.byref(n)), "loam_map_iter")
        return AtA, AtB, n.value

    def map_corr(self, n_cs, n_ss):
        a, b = np.empty((n_cs, 5), np.int32), np.empty((n_ss, 5), np.int32)
        self._check(self.lib.loam_map_get_corr(self._h, a.ctypes.data, n_cs, b.ctypes.data, n_ss), "loam_map_get_corr")
        return a, b

    def map_iter_partial(self, it, T, dev_ptr28):
        T = _f32(T)
        self._check(self.lib.loam_map_iter_partial(self._h, it, T.ctypes.data, dev_ptr28), "loam_map_iter_partial")


class LoamGpuPipeline:
    """Pipelined mode (loam_pipeline_*): submit sweeps, collect results in order; same results as LoamGpu.process_sweep."""

    def __init__(self, device=0, n_scans=16, ring_mode=0, ring_ang_min=-15.0, ring_ang_step=2.0, skip_frame_num=1,
                 want_registered=False, want_surround=False, max_points=None, max_map_points=None, pose_message_hop=False, gn_max_ctas=0):
        self.lib = load_library()
        p = Params()
        self.lib.loam_default_params(C.byref(p))
        p.n_scans, p.ring_mode, p.ring_ang_min, p.ring_ang_step = n_scans, ring_mode, ring_ang_min, ring_ang_step
        p.skip_frame_num = skip_frame_num
        p.want_registered, p.want_surround = int(want_registered), int(want_surround)
        p.pose_message_hop = int(pose_message_hop)
        p.gn_max_ctas = int(gn_max_ctas)
        if max_points:
            p.max_points = int(max_points)
        if max_map_points:
            p.max_map_points = int(max_map_points)
        self._h = C.c_void_p()
        rc = self.lib.loam_pipeline_create(C.byref(p), device, C.byref(self._h))
        if rc:
            raise LoamError(rc, "loam_pipeline_create", self.lib.loam_strerror(rc).decode())

    def _check(self, rc, where):
        if rc != 0:
            detail = self.lib.loam_strerror(rc).decode()
            if rc == LOAM_ECUDA:
                detail += ": " + self.lib.loam_pipeline_last_error(self._h).decode()
            raise LoamError(rc, where, detail)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.lib.loam_pipeline_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        self._check(self.lib.loam_pipeline_reset(self._h), "loam_pipeline_reset")

    def submit(self, xyz, stamp=0.0):
        xyz = _f32(xyz)
        self._check(self.lib.loam_pipeline_submit(self._h, xyz.ctypes.data, xyz.shape[0], xyz.strides[0] if xyz.shape[0] > 1 else 12, stamp),
                    "loam_pipeline_submit")

    def imu_push(self, stamp, quat_xyzw, angular_velocity, linear_acceleration):
        """loam_pipeline_imu_push: one /imu/data message, applied in submission order with the sweeps."""
        q, a, l = (np.ascontiguousarray(v, np.float64) for v in (quat_xyzw, angular_velocity, linear_acceleration))
        self._check(self.lib.loam_pipeline_imu_push(self._h, float(stamp), q.ctypes.data, a.ctypes.data, l.ctypes.data), "loam_pipeline_imu_push")

    def submit_device(self, dev_ptr, n, stride_bytes=12, stamp=0.0):
        self._check(self.lib.loam_pipeline_submit_device(self._h, dev_ptr, n, stride_bytes, stamp), "loam_pipeline_submit_device")

    def wait(self):
        r = SweepResult()
        self._check(self.lib.loam_pipeline_wait(self._h, C.byref(r)), "loam_pipeline_wait")
        return r

    @property
    def pending(self):
        return self.lib.loam_pipeline_pending(self._h)

    def stream(self, which):
        return self.lib.loam_pipeline_stream(self._h, which)

    def output_cloud(self, which="surround"):
        """Cloud held by the pipeline's output stage (loam_pipeline_handle(p, 3)): /laser_cloud_surround of the last run that
        published it (want_surround).  Read when the pipeline is idle."""
        h = self.lib.loam_pipeline_handle(self._h, 3)
        if not h:
            raise LoamError(-4, "loam_pipeline_handle", "created without want_surround")
        n = C.c_int()
        w = CLOUD[which] if isinstance(which, str) else which
        self._check(self.lib.loam_get_cloud(h, w, None, 0, C.byref(n)), "loam_get_cloud")
        out = np.empty((n.value, 4), np.float32)
        if n.value:
            self._check(self.lib.loam_get_cloud(h, w, out.ctypes.data, n.value, C.byref(n)), "loam_get_cloud")
        return out

    def stage_host_times(self, which, clear=True):
        """loam_host_times of the handle behind stage `which` (read when the pipeline is idle)."""
        a = (C.c_double * 16)()
        self._check(self.lib.loam_host_times(self.lib.loam_pipeline_handle(self._h, which), a, 1 if clear else 0), "loam_host_times")
        return list(a)

    def stage_times(self, clear=True):
        """Seconds of work per stage thread (extract, odometry, mapping) since the last clear; read when idle."""
        a = (C.c_double * 3)()
        self._check(self.lib.loam_pipeline_stage_times(self._h, a, 1 if clear else 0), "loam_pipeline_stage_times")
        return [a[0], a[1], a[2]]

    def stats(self):
        out = (C.c_longlong * 4)()
        self._check(self.lib.loam_pipeline_stats(self._h, out), "loam_pipeline_stats")
        return dict(launches=out[0], h2d_bytes=out[1], d2h_bytes=out[2], syncs=out[3])


def pipeline_submit_batch(pipes, sweeps, stride_bytes=12, lockstep=False, stamps=None):
    """loam_pipeline_submit_batch: one sweep per LoamGpuPipeline, the extraction of all of them in one batched launch chain;
    lockstep=True (loam_pipeline_submit_lockstep) batches the scan-to-scan odometry as well."""
    B = len(pipes)
    arrs = [_f32(x) for x in sweeps]
    ps = (C.c_void_p * B)(*[p._h for p in pipes])
    ptrs = (C.c_void_p * B)(*[a.ctypes.data for a in arrs])
    ns = (C.c_int * B)(*[a.shape[0] for a in arrs])
    fn = load_library().loam_pipeline_submit_lockstep if lockstep else load_library().loam_pipeline_submit_batch
    st = (C.c_double * B)(*[float(t) for t in stamps]) if stamps is not None else None
    rc = fn(ps, B, ptrs, ns, stride_bytes, st)
    if rc:
        raise LoamError(rc, "loam_pipeline_submit_lockstep" if lockstep else "loam_pipeline_submit_batch", load_library().loam_last_cuda_error(None).decode())


def odometry_process_batch(handles):
    """loam_odometry_process_batch: scan-to-scan odometry of the current sweep of every handle in lock-step (one launch per
    kernel and round for the whole batch).  Returns the list of OdomResult."""
    B = len(handles)
    hs = (C.c_void_p * B)(*[h._h for h in handles])
    out = (OdomResult * B)()
    rc = load_library().loam_odometry_process_batch(hs, B, out)
    if rc:
        raise LoamError(rc, "loam_odometry_process_batch", load_library().loam_last_cuda_error(None).decode())
    return list(out)


def extract_batch(handles, sweeps, stride_bytes=12):
    """loam_extract_batch: one sweep (float32 array (n, 3), or (n, k) rows of stride_bytes) per LoamGpu handle, every
    extraction kernel launched once for the whole batch.  Returns the list of Counts."""
    B = len(handles)
    arrs = [_f32(x) for x in sweeps]
    hs = (C.c_void_p * B)(*[h._h for h in handles])
    ptrs = (C.c_void_p * B)(*[a.ctypes.data for a in arrs])
    ns = (C.c_int * B)(*[a.shape[0] for a in arrs])
    out = (Counts * B)()
    rc = load_library().loam_extract_batch(hs, B, ptrs, ns, stride_bytes, None, out)
    if rc:
        raise LoamError(rc, "loam_extract_batch", load_library().loam_last_cuda_error(None).decode())
    return list(out)


def gn_solve(AtA, AtB, it, eig_threshold, state37):
    """Host-side 6x6 solve + degeneracy projection (LO:975-1004 / LM:968-997); state37 is updated in place."""
    lib = load_library()
    AtA, AtB = _f32(AtA), _f32(AtB)
    X = np.zeros(6, np.float32)
    rc = lib.loam_gn_solve(AtA.ctypes.data, AtB.ctypes.data, it, eig_threshold, state37.ctypes.data, X.ctypes.data)
    if rc:
        raise LoamError(rc, "loam_gn_solve")
    return X


def finish_reduced(reduced28):
    lib = load_library()
    r = np.ascontiguousarray(reduced28, np.float64)
    AtA, AtB, n = np.zeros((6, 6), np.float32), np.zeros(6, np.float32), C.c_int()
    rc = lib.loam_map_finish_reduced(r.ctypes.data, AtA.ctypes.data, AtB.ctypes.data, C.byref(n))
    if rc:
        raise LoamError(rc, "loam_map_finish_reduced")
    return AtA, AtB, n.value


# ------------------------------------------------------------------------------------------------ N4 track calibration
def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class TrackCalibration:
    """Mirror of the reference's trackCalibration class (track_calibration.h:12-65): constructed from a SLAM track, the
    matching ENU (GPS) track and per-point weights, then do_icp() and do_calibration() like SD:241-243 / LD:67-72.
    mode 0 runs the smoothing loop TC:648-674 on the GPU in the serial loop's order, mode 1 its closed form on the host."""

    def __init__(self, slam_xyzt, enu_xyzt, weights, mode=0, device=0):
        self.slam, self.enu, self.w = _f64(slam_xyzt), _f64(enu_xyzt), _f64(weights)
        if not (self.slam.shape == self.enu.shape and self.slam.shape[0] == self.w.shape[0] and self.slam.shape[1] == 4):
            raise ValueError("there's something wrong in icp data no, please check it out")  # TC:46-50 (the reference exits)
        self.n, self.mode, self.device = self.slam.shape[0], mode, device
        self.T = None
        self.rotated = None

    def do_icp(self):
        self.T = np.zeros((4, 4))
        self.rotated = np.zeros((self.n, 2))
        rc = load_library().loam_track_icp(self.slam.ctypes.data, self.enu.ctypes.data, self.w.ctypes.data, self.n, self.T.ctypes.data,
                                           self.rotated.ctypes.data)
        if rc:
            raise LoamError(rc, "loam_track_icp")
        return 1

    def do_calibration(self):
        if self.rotated is None:
            raise LoamError(-4, "loam_track_smooth", "do_icp() first")
        out = np.zeros((self.n, 4))
        rc = load_library().loam_track_smooth(self.rotated.ctypes.data, self.enu.ctypes.data, self.n, self.mode, self.device, out.ctypes.data)
        if rc:
            raise LoamError(rc, "loam_track_smooth", load_library().loam_last_cuda_error(None).decode())
        return out


def track_speed_weights(slam_xyzt):
    s = _f64(slam_xyzt)
    w = np.zeros(s.shape[0])
    rc = load_library().loam_track_speed_weights(s.ctypes.data, s.shape[0], w.ctypes.data)
    if rc:
        raise LoamError(rc, "loam_track_speed_weights")
    return w


def track_residual_weights(slam_xyzt, enu_xyzt, cal_xyzt):
    s, e, c = _f64(slam_xyzt), _f64(enu_xyzt), _f64(cal_xyzt)
    w = np.zeros(s.shape[0])
    rc = load_library().loam_track_residual_weights(s.ctypes.data, e.ctypes.data, c.ctypes.data, s.shape[0], w.ctypes.data)
    if rc:
        raise LoamError(rc, "loam_track_residual_weights")
    return w


def track_calibrate_long(slam_xyzt, enu_xyzt, iterations=5, mode=0, device=0):
    """LD:57-83: returns (weights merged into /gps_weight, last calibrated track)."""
    s, e = _f64(slam_xyzt), _f64(enu_xyzt)
    n = s.shape[0]
    w, cal = np.zeros(n), np.zeros((n, 4))
    rc = load_library().loam_track_calibrate_long(s.ctypes.data, e.ctypes.data, n, iterations, mode, device, w.ctypes.data, cal.ctypes.data)
    if rc:
        raise LoamError(rc, "loam_track_calibrate_long", load_library().loam_last_cuda_error(None).decode())
    return w, cal


def track_svd3(H):
    h = _f64(H).reshape(3, 3)
    U, S, V = np.zeros((3, 3)), np.zeros(3), np.zeros((3, 3))
    rc = load_library().loam_track_svd3(h.ctypes.data, U.ctypes.data, S.ctypes.data, V.ctypes.data)
    if rc:
        raise LoamError(rc, "loam_track_svd3")
    return U, S, V
