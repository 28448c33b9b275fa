"""ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes binding of oracle/liborc.so (the CPU restatement of the reference).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class PipelineResult(C.Structure):
    _fields_ = [("odom", C.c_float * 6), ("mapped", C.c_float * 6), ("rel", C.c_float * 6), ("odom_published", C.c_int),
                ("mapping_ran", C.c_int), ("odom_iters", C.c_int), ("map_iters", C.c_int), ("n_full", C.c_int),
                ("n_sharp", C.c_int), ("n_less_sharp", C.c_int), ("n_flat", C.c_int), ("n_less_flat", C.c_int),
                ("n_corner_stack", C.c_int), ("n_surf_stack", C.c_int), ("n_corner_map", C.c_int), ("n_surf_map", C.c_int),
                ("t_extract", C.c_double), ("t_odom", C.c_double), ("t_map", C.c_double)]


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "liborc.so"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liborc.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        vp, ip = C.c_void_p, C.POINTER(C.c_int)
        L.orc_voxel_grid.argtypes = [vp, C.c_int, C.c_float, vp, C.c_int]
        L.orc_knn.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, vp, vp]
        L.orc_gemm.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int]
        L.orc_qr_solve.argtypes = [vp, vp, vp, C.c_int, C.c_int]
        L.orc_jacobi_eigen.argtypes = [vp, vp, vp, C.c_int]
        L.orc_lu_inverse.argtypes = [vp, vp, C.c_int]
        L.orc_svd3.argtypes = [vp, vp, vp, vp]
        L.orc_gn_solve.argtypes = [vp, vp, C.c_int, C.c_float, vp, vp]
        L.orc_sr_create.restype = vp
        L.orc_sr_create.argtypes = [C.c_int, C.c_int, C.c_float, C.c_float]
        L.orc_sr_destroy.argtypes = [vp]
        L.orc_sr_extract.argtypes = [vp, vp, C.c_int, C.c_int]
        L.orc_sr_imu.argtypes = [vp, C.c_double, vp, vp, vp]
        L.orc_sr_extract_imu.argtypes = [vp, vp, C.c_int, C.c_int, C.c_double, vp]
        L.orc_sr_cloud.argtypes = [vp, C.c_int, vp, C.c_int]
        L.orc_lonode_create.restype = vp
        L.orc_lonode_destroy.argtypes = [vp]
        L.orc_lonode_control.argtypes = [vp, C.c_int]
        L.orc_lonode_step.argtypes = [vp, vp, C.c_int, vp, C.c_int, vp, C.c_int, vp, C.c_int, vp, C.c_int, vp, vp]
        L.orc_lonode_cloud.argtypes = [vp, C.c_int, vp, C.c_int]
        L.orc_sr_ints.argtypes = [vp, C.c_int, vp, C.c_int]
        L.orc_sr_curvature.argtypes = [vp, vp, C.c_int]
        L.orc_transform_to_start.argtypes = [vp, C.c_int, vp, vp]
        L.orc_transform_to_end.argtypes = [vp, C.c_int, vp, vp, vp]
        L.orc_odom_iteration.argtypes = [vp, C.c_int, vp, C.c_int, vp, C.c_int, vp, C.c_int, vp, C.c_int, C.c_int, vp, vp, vp,
                                         vp, vp, vp, vp, ip, vp, vp, C.c_int]
        L.orc_associate_to_map.argtypes = [vp, C.c_int, vp, vp]
        L.orc_associate_tobe_mapped.argtypes = [vp, C.c_int, vp, vp]
        L.orc_transform_associate_to_map.argtypes = [vp, vp, vp, vp, vp]
        L.orc_map_iteration.argtypes = [vp, C.c_int, vp, C.c_int, vp, C.c_int, vp, C.c_int, vp, C.c_int, vp, vp, vp, vp, ip]
        L.orc_map_iteration_sums28.argtypes = [vp, C.c_int, vp, C.c_int, vp, C.c_int, vp, C.c_int, vp, vp]
        L.orc_lm_create.restype = vp
        L.orc_lm_create.argtypes = [C.c_int]
        L.orc_lm_destroy.argtypes = [vp]
        L.orc_lm_step.argtypes = [vp, vp, C.c_int, vp, C.c_int, vp, C.c_int, vp, C.c_int, vp]
        L.orc_tm_create.restype = vp
        L.orc_tm_destroy.argtypes = [vp]
        L.orc_tm_odometry.argtypes = [vp, vp, C.c_double, vp, vp]
        L.orc_tm_aft_mapped.argtypes = [vp, vp, vp]
        L.orc_pipeline_create.restype = vp
        L.orc_pipeline_create.argtypes = [C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int]
        L.orc_pipeline_destroy.argtypes = [vp]
        L.orc_pipeline_reset.argtypes = [vp]
        L.orc_pipeline_bef_mapped.argtypes = [vp, vp]
        L.orc_pipeline_set_ros_hop.argtypes = [vp, C.c_int]
        L.orc_odometry_ros_hop.argtypes = [vp, vp]
        L.orc_pipeline_process.argtypes = [vp, vp, C.c_int, C.c_int, C.POINTER(PipelineResult)]
        L.orc_pipeline_cloud.argtypes = [vp, C.c_int, vp, C.c_int]
        L.orc_pipeline_run_threaded.argtypes = [vp, vp, vp, C.c_int, vp]
        L.orc_pipeline_map_size.argtypes = [vp, ip, ip]
        _LIB = L
    return _LIB


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def voxel_grid(pts4, leaf):
    pts4 = _f32(pts4)
    out = np.empty_like(pts4)
    v = lib().orc_voxel_grid(pts4.ctypes.data, pts4.shape[0], leaf, out.ctypes.data, out.shape[0])
    return out[:v].copy()


def knn(cloud4, q4, k, brute=False):
    cloud4, q4 = _f32(cloud4), _f32(q4)
    idx = np.empty((q4.shape[0], k), np.int32)
    d2 = np.empty((q4.shape[0], k), np.float32)
    lib().orc_knn(cloud4.ctypes.data, cloud4.shape[0], q4.ctypes.data, q4.shape[0], k, int(brute), idx.ctypes.data, d2.ctypes.data)
    return idx, d2


def gemm(A, B):
    A, B = _f32(A), _f32(B)
    Cm = np.empty((A.shape[0], B.shape[1]), np.float32)
    lib().orc_gemm(A.ctypes.data, B.ctypes.data, Cm.ctypes.data, A.shape[0], A.shape[1], B.shape[1])
    return Cm


def qr_solve(A, b):
    A, b = _f32(A), _f32(b)
    x = np.zeros(A.shape[1], np.float32)
    lib().orc_qr_solve(A.ctypes.data, b.ctypes.data, x.ctypes.data, A.shape[0], A.shape[1])
    return x


def jacobi_eigen(A):
    A = _f32(A)
    n = A.shape[0]
    W, V = np.zeros(n, np.float32), np.zeros((n, n), np.float32)
    lib().orc_jacobi_eigen(A.ctypes.data, W.ctypes.data, V.ctypes.data, n)
    return W, V


def lu_inverse(A):
    A = _f32(A)
    out = np.zeros_like(A)
    lib().orc_lu_inverse(A.ctypes.data, out.ctypes.data, A.shape[0])
    return out


def svd3(H):
    """Restated Eigen::JacobiSVD of a 3x3 fp64 matrix (orc_linalg.h svd3_jacobi): H = U diag(S) V^T."""
    h = np.ascontiguousarray(H, np.float64).reshape(3, 3)
    U, S, V = np.zeros((3, 3)), np.zeros(3), np.zeros((3, 3))
    lib().orc_svd3(h.ctypes.data, U.ctypes.data, S.ctypes.data, V.ctypes.data)
    return U, S, V


def gn_solve(AtA, AtB, it, thre, state37):
    AtA, AtB = _f32(AtA), _f32(AtB)
    X = np.zeros(6, np.float32)
    lib().orc_gn_solve(AtA.ctypes.data, AtB.ctypes.data, it, thre, state37.ctypes.data, X.ctypes.data)
    return X


class ScanRegistration:
    CLOUDS = ("full", "sharp", "less_sharp", "flat", "less_flat")

    def __init__(self, n_scans=16, ring_mode=0, ang_min=-15.0, ang_step=2.0):
        self._h = lib().orc_sr_create(n_scans, ring_mode, ang_min, ang_step)
        self.n_scans = n_scans

    def __del__(self):
        try:
            lib().orc_sr_destroy(self._h)
        except Exception:
            pass

    def extract(self, xyz):
        xyz = _f32(xyz)
        lib().orc_sr_extract(self._h, xyz.ctypes.data, xyz.shape[0], xyz.strides[0] // 4)
        return {k: self.cloud(k) for k in self.CLOUDS}

    def imu(self, stamp, quat_xyzw, angular_velocity, linear_acceleration):
        """One /imu/data message (imuHandler SR:754-837)."""
        q, a, l = (np.ascontiguousarray(v, np.float64) for v in (quat_xyzw, angular_velocity, linear_acceleration))
        lib().orc_sr_imu(self._h, float(stamp), q.ctypes.data, a.ctypes.data, l.ctypes.data)

    def extract_imu(self, xyz, stamp):
        """laserCloudHandler with the IMU branch (SR:364-434): returns (clouds, the 12 floats of /imu_trans)."""
        xyz = _f32(xyz)
        tr = np.zeros(12, np.float32)
        lib().orc_sr_extract_imu(self._h, xyz.ctypes.data, xyz.shape[0], xyz.strides[0] // 4, float(stamp), tr.ctypes.data)
        return {k: self.cloud(k) for k in self.CLOUDS}, tr

    def cloud(self, which):
        w = self.CLOUDS.index(which)
        n = lib().orc_sr_cloud(self._h, w, None, 0)
        out = np.empty((n, 4), np.float32)
        lib().orc_sr_cloud(self._h, w, out.ctypes.data, n)
        return out

    def ints(self, which):
        w = ("scan_start", "scan_end", "picked_mask", "label", "sort_ind").index(which)
        n = lib().orc_sr_ints(self._h, w, None, 0)
        out = np.empty(n, np.int32)
        lib().orc_sr_ints(self._h, w, out.ctypes.data, n)
        return out

    def curvature(self):
        n = lib().orc_sr_curvature(self._h, None, 0)
        out = np.empty(n, np.float32)
        lib().orc_sr_curvature(self._h, out.ctypes.data, n)
        return out


class LaserOdometry:
    """The odometry node alone: one main-loop body (LO:502-1147) per step, /imu_trans as an input."""

    def __init__(self):
        self._h = lib().orc_lonode_create()

    def __del__(self):
        try:
            lib().orc_lonode_destroy(self._h)
        except Exception:
            pass

    def control(self, inited):
        lib().orc_lonode_control(self._h, int(inited))

    def step(self, feat, imu12=None):
        """feat = [full, sharp, less_sharp, flat, less_flat]; returns (out18, [corner_last, surf_last, full_res] or None)."""
        full, sharp, less_sharp, flat, less_flat = (_f32(x) for x in feat)
        out = np.zeros(18, np.float32)
        imu = _f32(imu12).ctypes.data if imu12 is not None else None
        lib().orc_lonode_step(self._h, sharp.ctypes.data, sharp.shape[0], less_sharp.ctypes.data, less_sharp.shape[0], flat.ctypes.data,
                          flat.shape[0], less_flat.ctypes.data, less_flat.shape[0], full.ctypes.data, full.shape[0], imu, out.ctypes.data)
        clouds = None
        if out[13] > 0:
            clouds = []
            for w in range(3):
                n = lib().orc_lonode_cloud(self._h, w, None, 0)
                c = np.empty((n, 4), np.float32)
                if n:
                    lib().orc_lonode_cloud(self._h, w, c.ctypes.data, n)
                clouds.append(c)
        return out, clouds


def transform_to_end(pts4, T, imu12=None):
    pts4, T = _f32(pts4), _f32(T)
    out = np.empty_like(pts4)
    imu = _f32(imu12).ctypes.data if imu12 is not None else None
    lib().orc_transform_to_end(pts4.ctypes.data, pts4.shape[0], T.ctypes.data, imu, out.ctypes.data)
    return out


def transform_to_start(pts4, T):
    pts4, T = _f32(pts4), _f32(T)
    out = np.empty_like(pts4)
    lib().orc_transform_to_start(pts4.ctypes.data, pts4.shape[0], T.ctypes.data, out.ctypes.data)
    return out


class OdomIter:
    """Stateful wrapper over orc_odom_iteration (keeps the correspondence arrays between iterations)."""

    def __init__(self, sharp, flat, corner_last, surf_last, brute=False):
        self.sharp, self.flat, self.cl, self.sl = _f32(sharp), _f32(flat), _f32(corner_last), _f32(surf_last)
        self.brute = brute
        ns, nf = self.sharp.shape[0], self.flat.shape[0]
        self.c1, self.c2 = -np.ones(ns, np.int32), -np.ones(ns, np.int32)
        self.s1, self.s2, self.s3 = -np.ones(nf, np.int32), -np.ones(nf, np.int32), -np.ones(nf, np.int32)

    def iterate(self, it, T, want_rows=False):
        T = _f32(T)
        AtA, AtB, n = np.zeros((6, 6), np.float32), np.zeros(6, np.float32), C.c_int()
        cap = self.sharp.shape[0] + self.flat.shape[0]
        rowsA = np.zeros((cap, 6), np.float32) if want_rows else None
        rowsB = np.zeros(cap, np.float32) if want_rows else None
        lib().orc_odom_iteration(self.sharp.ctypes.data, self.sharp.shape[0], self.flat.ctypes.data, self.flat.shape[0],
                                 self.cl.ctypes.data, self.cl.shape[0], self.sl.ctypes.data, self.sl.shape[0], T.ctypes.data, it,
                                 int(self.brute), self.c1.ctypes.data, self.c2.ctypes.data, self.s1.ctypes.data,
                                 self.s2.ctypes.data, self.s3.ctypes.data, AtA.ctypes.data, AtB.ctypes.data, C.byref(n),
                                 rowsA.ctypes.data if want_rows else None, rowsB.ctypes.data if want_rows else None,
                                 cap if want_rows else 0)
        if want_rows:
            return AtA, AtB, n.value, rowsA[:n.value], rowsB[:n.value]
        return AtA, AtB, n.value


def associate_to_map(pts4, T):
    """pointAssociateToMap (LM:244-262) of a cloud."""
    a, T = _f32(pts4), _f32(T)
    out = np.empty_like(a)
    lib().orc_associate_to_map(a.ctypes.data, a.shape[0], T.ctypes.data, out.ctypes.data)
    return out


def map_iteration(corner_stack, surf_stack, corner_map, surf_map, T, brute=False):
    a, b, c, d, T = _f32(corner_stack), _f32(surf_stack), _f32(corner_map), _f32(surf_map), _f32(T)
    cc, cs = np.empty((a.shape[0], 5), np.int32), np.empty((b.shape[0], 5), np.int32)
    AtA, AtB, n = np.zeros((6, 6), np.float32), np.zeros(6, np.float32), C.c_int()
    lib().orc_map_iteration(a.ctypes.data, a.shape[0], b.ctypes.data, b.shape[0], c.ctypes.data, c.shape[0], d.ctypes.data,
                            d.shape[0], T.ctypes.data, int(brute), cc.ctypes.data, cs.ctypes.data, AtA.ctypes.data,
                            AtB.ctypes.data, C.byref(n))
    return AtA, AtB, n.value, cc, cs


def map_iteration_sums28(corner_stack, surf_stack, corner_map, surf_map, T):
    a, b, c, d, T = _f32(corner_stack), _f32(surf_stack), _f32(corner_map), _f32(surf_map), _f32(T)
    out = np.zeros(28, np.float64)
    lib().orc_map_iteration_sums28(a.ctypes.data, a.shape[0], b.ctypes.data, b.shape[0], c.ctypes.data, c.shape[0], d.ctypes.data,
                                   d.shape[0], T.ctypes.data, out.ctypes.data)
    return out


class LaserMapping:
    """Oracle laserMapping node driven directly (odometry message + optional full message set)."""

    def __init__(self, brute=False):
        self._h = lib().orc_lm_create(int(brute))

    def __del__(self):
        try:
            lib().orc_lm_destroy(self._h)
        except Exception:
            pass

    def step(self, Tsum, corner=None, surf=None, full=None):
        T = _f32(Tsum)
        out = np.zeros(25, np.float32)
        if corner is None:
            lib().orc_lm_step(self._h, T.ctypes.data, 0, None, 0, None, 0, None, 0, out.ctypes.data)
            return None
        c, s = _f32(corner), _f32(surf)
        f = _f32(full) if full is not None else np.zeros((0, 4), np.float32)
        lib().orc_lm_step(self._h, T.ctypes.data, 1, c.ctypes.data, c.shape[0], s.ctypes.data, s.shape[0], f.ctypes.data, f.shape[0],
                          out.ctypes.data)
        return out


class TransformMaintenance:
    """Oracle restatement of transformMaintenance.cpp (TM:116-157, 262-338)."""

    def __init__(self):
        self._h = lib().orc_tm_create()

    def __del__(self):
        try:
            lib().orc_tm_destroy(self._h)
        except Exception:
            pass

    def odometry(self, Tsum, stamp):
        T = _f32(Tsum)
        out, track = np.zeros(6, np.float32), np.zeros(4, np.float64)
        lib().orc_tm_odometry(self._h, T.ctypes.data, float(stamp), out.ctypes.data, track.ctypes.data)
        return out, track

    def aft_mapped(self, aft, bef):
        a, b = _f32(aft), _f32(bef)
        lib().orc_tm_aft_mapped(self._h, a.ctypes.data, b.ctypes.data)


class Pipeline:
    CLOUDS = ("full", "sharp", "less_sharp", "flat", "less_flat", "corner_last", "surf_last", "full_res3", "surround", "registered")

    def __init__(self, n_scans=16, ring_mode=0, ang_min=-15.0, ang_step=2.0, brute=False, keep_clouds=True):
        self._h = lib().orc_pipeline_create(n_scans, ring_mode, ang_min, ang_step, int(brute), int(keep_clouds))

    def __del__(self):
        try:
            lib().orc_pipeline_destroy(self._h)
        except Exception:
            pass

    def reset(self):
        lib().orc_pipeline_reset(self._h)

    def set_ros_hop(self, on=True):
        """Route the odometry pose through the quaternion message like the reference's ROS nodes do (LO:1066-1078, LM:322-332)."""
        lib().orc_pipeline_set_ros_hop(self._h, int(on))

    def process(self, xyz):
        xyz = _f32(xyz)
        r = PipelineResult()
        lib().orc_pipeline_process(self._h, xyz.ctypes.data, xyz.shape[0], xyz.strides[0] // 4, C.byref(r))
        return r

    def bef_mapped(self):
        out = np.zeros(6, np.float32)
        lib().orc_pipeline_bef_mapped(self._h, out.ctypes.data)
        return out

    def run_threaded(self, xyz_all, offsets):
        """Three stage threads (SR | LO | LM) over a whole sequence; returns the per-sweep PipelineResult array."""
        xyz_all = _f32(xyz_all)
        offsets = np.ascontiguousarray(offsets, np.int64)
        n = offsets.shape[0] - 1
        res = (PipelineResult * n)()
        lib().orc_pipeline_run_threaded(self._h, xyz_all.ctypes.data, offsets.ctypes.data, n, res)
        return res

    def cloud(self, which):
        w = self.CLOUDS.index(which)
        n = lib().orc_pipeline_cloud(self._h, w, None, 0)
        out = np.empty((n, 4), np.float32)
        lib().orc_pipeline_cloud(self._h, w, out.ctypes.data, n)
        return out

    def map_size(self):
        a, b = C.c_int(), C.c_int()
        lib().orc_pipeline_map_size(self._h, C.byref(a), C.byref(b))
        return a.value, b.value
