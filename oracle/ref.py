"""ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes driver of oracle/_ref/*.so: the reference's OWN scanRegistration.cpp,
laserOdometry.cpp and laserMapping.cpp, compiled unmodified from /root/reference against the shim headers in
oracle/ref_build/shim (recipe: oracle/ref_build/Makefile).  The three nodes are wired the way gpsCalibration.launch
wires them: clouds and the odometry message (position + quaternion) are passed from node to node.

Each node's file-scope state lives in its shared object, so there is ONE pipeline per process.
"""
import atexit
import ctypes as C
import os
import queue
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_DIR = os.path.join(_HERE, "_ref")
_LIBS = None


def available():
    return all(os.path.exists(os.path.join(_DIR, f)) for f in ("libref_sr.so", "libref_lo.so", "libref_lm.so"))


def _libs():
    global _LIBS
    if _LIBS is None:
        vp, ip = C.c_void_p, C.POINTER(C.c_int)
        sr = C.CDLL(os.path.join(_DIR, "libref_sr.so"))
        lo = C.CDLL(os.path.join(_DIR, "libref_lo.so"))
        lm = C.CDLL(os.path.join(_DIR, "libref_lm.so"))
        sr.ref_sr_process.argtypes = [vp, C.c_int, C.c_double]
        sr.ref_sr_cloud.argtypes = [C.c_int, vp, C.c_int]
        lo.ref_lo_step.argtypes = [vp, C.c_int, vp, C.c_int, vp, C.c_int, vp, C.c_int, vp, C.c_int, vp, C.c_double, vp]
        lo.ref_lo_control.argtypes = [C.c_int]
        lo.ref_lo_odometry.argtypes = [vp]
        lo.ref_lo_cloud.argtypes = [C.c_int, vp, C.c_int]
        lm.ref_lm_odometry_only.argtypes = [vp, C.c_double]
        lm.ref_lm_step.argtypes = [vp, C.c_int, vp, C.c_int, vp, C.c_int, vp, C.c_double, vp]
        lm.ref_lm_cloud.argtypes = [C.c_int, vp, C.c_int]
        lm.ref_lm_map_size.argtypes = [ip, ip]
        lo.ref_lo_start()
        lm.ref_lm_start()
        _LIBS = (sr, lo, lm)
        atexit.register(shutdown)
    return _LIBS


_TM = None


def pose7_from_T(T):
    """The nav_msgs/Odometry pose the nodes publish for a pose vector T (LO:1066-1078, LM:1114-1124): position +
    quaternion built from createQuaternionMsgFromRollPitchYaw(rz, -rx, -ry) with the axes permuted."""
    T = np.asarray(T, np.float32)
    roll, pitch, yaw = float(T[2]), float(-T[0]), float(-T[1])
    hy, hp, hr = yaw * 0.5, pitch * 0.5, roll * 0.5
    cy, sy, cp, sp, cr, sr = np.cos(hy), np.sin(hy), np.cos(hp), np.sin(hp), np.cos(hr), np.sin(hr)
    gx = sr * cp * cy - cr * sp * sy
    gy = cr * sp * cy + sr * cp * sy
    gz = cr * cp * sy - sr * sp * cy
    gw = cr * cp * cy + sr * sp * sy
    return np.array([float(T[3]), float(T[4]), float(T[5]), -gy, -gz, gx, gw], np.float64)


def tm():
    """The reference's transformMaintenance node (oracle/_ref/libref_tm.so)."""
    global _TM
    if _TM is None:
        vp = C.c_void_p
        L = C.CDLL(os.path.join(_DIR, "libref_tm.so"))
        L.ref_tm_odometry.argtypes = [vp, C.c_double, vp, vp]
        L.ref_tm_aft_mapped.argtypes = [vp, vp, C.c_double]
        L.ref_tm_start()
        atexit.register(L.ref_tm_stop)
        _TM = L
    return _TM


def tm_odometry(Tsum, stamp):
    p = pose7_from_T(Tsum)
    out, track = np.zeros(6, np.float32), np.zeros(4, np.float64)
    tm().ref_tm_odometry(p.ctypes.data, float(stamp), out.ctypes.data, track.ctypes.data)
    return out, track


def tm_aft_mapped(aft, bef, stamp):
    p = pose7_from_T(aft)
    b = np.ascontiguousarray(bef, np.float32)
    tm().ref_tm_aft_mapped(p.ctypes.data, b.ctypes.data, float(stamp))


def shutdown():
    """Stops the two node loop threads (must run before the interpreter exits)."""
    global _LIBS
    if _LIBS is not None:
        _LIBS[1].ref_lo_stop()
        _LIBS[2].ref_lm_stop()
        _LIBS = None


def _cloud(fn, which):
    n = fn(which, None, 0)
    out = np.empty((n, 4), np.float32)
    if n:
        fn(which, out.ctypes.data, n)
    return out


class SweepOut:
    __slots__ = ("features", "odom", "rel", "odom_published", "clouds_published", "fullres_published", "mapping_ran", "mapped",
                 "bef_mapped", "tobe_mapped")


def sr_process(xyz, stamp):
    sr, _, _ = _libs()
    xyz = np.ascontiguousarray(xyz, np.float32)
    sr.ref_sr_process(xyz.ctypes.data, xyz.shape[0], float(stamp))
    return [_cloud(sr.ref_sr_cloud, w) for w in range(5)]  # full, sharp, less sharp, flat, less flat


def lo_step(feat, stamp):
    _, lo, _ = _libs()
    full, sharp, less_sharp, flat, less_flat = feat
    out = np.zeros(18, np.float32)
    lo.ref_lo_step(sharp.ctypes.data, sharp.shape[0], less_sharp.ctypes.data, less_sharp.shape[0], flat.ctypes.data, flat.shape[0],
                   less_flat.ctypes.data, less_flat.shape[0], full.ctypes.data, full.shape[0], None, float(stamp), out.ctypes.data)
    pose7 = np.zeros(7, np.float64)
    lo.ref_lo_odometry(pose7.ctypes.data)
    clouds = None
    if out[13] > 0:
        clouds = [_cloud(lo.ref_lo_cloud, w) for w in range(3)]
    return out, pose7, clouds


def lm_step(clouds, pose7, stamp, full_set):
    _, _, lm = _libs()
    if not full_set:
        lm.ref_lm_odometry_only(pose7.ctypes.data, float(stamp))
        return None
    out = np.zeros(24, np.float32)
    c, s, f = clouds
    lm.ref_lm_step(c.ctypes.data, c.shape[0], s.ctypes.data, s.shape[0], f.ctypes.data, f.shape[0], pose7.ctypes.data, float(stamp),
                   out.ctypes.data)
    return out


def control_reset():
    """IMControl{systemInited=false} to laserOdometry (IN:281-284)."""
    _libs()[1].ref_lo_control(0)


def process(xyz, stamp):
    """One sweep through the three reference nodes, sequentially."""
    r = SweepOut()
    feat = sr_process(xyz, stamp)
    r.features = feat
    o, pose7, clouds = lo_step(feat, stamp)
    r.odom, r.rel = o[:6].copy(), o[6:12].copy()
    r.odom_published, r.clouds_published, r.fullres_published = bool(o[12] > 0), bool(o[13] > 0), bool(o[14] > 0)
    r.mapping_ran = False
    r.mapped = r.bef_mapped = r.tobe_mapped = None
    if r.odom_published:
        m = lm_step(clouds, pose7, stamp, r.fullres_published)
        if m is not None:
            r.mapping_ran = True
            r.mapped, r.bef_mapped, r.tobe_mapped = m[:6].copy(), m[6:12].copy(), m[12:18].copy()
    return r


def map_size():
    a, b = C.c_int(), C.c_int()
    _libs()[2].ref_lm_map_size(C.byref(a), C.byref(b))
    return a.value, b.value


def run_sequence(arr, offs, t0=100.0):
    """A whole sequence with the three nodes running concurrently (three Python threads; ctypes releases the GIL),
    which is how the reference runs them as three ROS processes.  Returns the list of final poses per sweep."""
    control_reset()
    n = len(offs) - 1
    q1, q2 = queue.Queue(maxsize=8), queue.Queue(maxsize=8)
    res = [None] * n

    def a():
        for k in range(n):
            q1.put((k, sr_process(arr[offs[k]:offs[k + 1]], t0 + 0.1 * k)))
        q1.put(None)

    def b():
        while True:
            it = q1.get()
            if it is None:
                q2.put(None)
                return
            k, feat = it
            o, pose7, clouds = lo_step(feat, t0 + 0.1 * k)
            q2.put((k, o, pose7, clouds))

    def c():
        while True:
            it = q2.get()
            if it is None:
                return
            k, o, pose7, clouds = it
            m = None
            if o[12] > 0:
                m = lm_step(clouds, pose7, t0 + 0.1 * k, o[14] > 0)
            res[k] = (o[:6].copy(), None if m is None else m[:6].copy())

    ts = [threading.Thread(target=f) for f in (a, b, c)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    return res


# ---------------------------------------------------------------------------------------------- input_data.cpp (N1)
_PUBLISH = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int))
_CONTROL = C.CFUNCTYPE(None, C.c_void_p)
_TRACK = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.POINTER(C.c_double), C.c_int)


class _ReplayCallbacks(C.Structure):
    _fields_ = [("publish", _PUBLISH), ("control", _CONTROL), ("slam_track", _TRACK), ("user", C.c_void_p)]


def input_data_available():
    return os.path.exists(os.path.join(_DIR, "libref_in.so"))


def input_replay(messages_per_bag, stamps, long_distance, short_distance, overlap_distance, publish, control, slam_track):
    """Runs the reference's own input_data.cpp main() once over in-memory bags.  publish(bag, msg) -> (stamp, odometry
    or None); control(); slam_track(flag, ndarray (n, 4)).  The node keeps its state in file-scope globals, so every
    run loads a private copy of the shared object."""
    import shutil
    import tempfile
    tmp = tempfile.mkdtemp(prefix="ref_in_")
    try:
        so = os.path.join(tmp, "libref_in_run.so")
        shutil.copy(os.path.join(_DIR, "libref_in.so"), so)
        L = C.CDLL(so)
        L.ref_in_run.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double, C.POINTER(_ReplayCallbacks), C.c_char_p]
        errors = []

        def _pub(user, bag, msg, stamp, odo, arrived):
            try:
                s, o = publish(bag, msg)
                stamp[0] = s
                arrived[0] = 0 if o is None else 1
                if o is not None:
                    for i in range(4):
                        odo[i] = o[i]
            except BaseException as e:
                errors.append(e)
                arrived[0] = 0

        def _ctl(user):
            try:
                control()
            except BaseException as e:
                errors.append(e)

        def _trk(user, flag, xyzt, n):
            slam_track(flag, np.ctypeslib.as_array(xyzt, shape=(n, 4)).copy() if n else np.zeros((0, 4)))

        cb = _ReplayCallbacks(_PUBLISH(_pub), _CONTROL(_ctl), _TRACK(_trk), None)
        counts = np.ascontiguousarray(messages_per_bag, np.int32)
        flat = np.ascontiguousarray(np.concatenate([np.asarray(s, np.float64) for s in stamps]) if len(stamps) else np.zeros(0))
        rc = L.ref_in_run(counts.ctypes.data, flat.ctypes.data, len(messages_per_bag), long_distance, short_distance, overlap_distance,
                          C.byref(cb), os.path.join(tmp, "baglist.txt").encode())
        if errors:
            raise errors[0]
        return rc
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


# ------------------------------------------------------------------------------------------------ N4 track calibration
_TC = None


def tc_available():
    return os.path.exists(os.path.join(_DIR, "libref_tc.so"))


def _tc():
    """The reference's own track_calibration.cc + weight_calculation.cc (oracle/ref_build/ref_tc.cpp)."""
    global _TC
    if _TC is None:
        L = C.CDLL(os.path.join(_DIR, "libref_tc.so"))
        vp = C.c_void_p
        L.ref_wc_speed.argtypes = [vp, C.c_int, vp]
        L.ref_wc_residual.argtypes = [vp, vp, vp, C.c_int, vp]
        L.ref_tc_calibrate.argtypes = [vp, vp, vp, C.c_int, vp, vp]
        L.ref_ld_long.argtypes = [vp, vp, C.c_int, C.c_int, vp, vp]
        _TC = L
    return _TC


def _d(a):
    return np.ascontiguousarray(a, np.float64)


def tc_speed_weights(slam):
    s = _d(slam)
    w = np.zeros(s.shape[0])
    _tc().ref_wc_speed(s.ctypes.data, s.shape[0], w.ctypes.data)
    return w


def tc_residual_weights(slam, enu, cal):
    s, e, c = _d(slam), _d(enu), _d(cal)
    w = np.zeros(s.shape[0])
    _tc().ref_wc_residual(s.ctypes.data, e.ctypes.data, c.ctypes.data, s.shape[0], w.ctypes.data)
    return w


def tc_calibrate(slam, enu, w):
    """(calibrated track n x 4, SLAMRotatedCoord n x 2) of one trackCalibration object."""
    s, e, ww = _d(slam), _d(enu), _d(w)
    n = s.shape[0]
    cal, rot = np.zeros((n, 4)), np.zeros((n, 2))
    _tc().ref_tc_calibrate(s.ctypes.data, e.ctypes.data, ww.ctypes.data, n, cal.ctypes.data, rot.ctypes.data)
    return cal, rot


def tc_long(slam, enu, iterations=5):
    s, e = _d(slam), _d(enu)
    n = s.shape[0]
    w, cal = np.zeros(n), np.zeros((n, 4))
    _tc().ref_ld_long(s.ctypes.data, e.ctypes.data, n, iterations, w.ctypes.data, cal.ctypes.data)
    return w, cal


# ------------------------------------------------------------------------------------------------ SR with the IMU branch
class SrWithImu:
    """A PRIVATE copy of the reference's scanRegistration.cpp (its IMU state lives in file-scope globals and would leak into
    every later sweep of the process): imuHandler (SR:754-837) + laserCloudHandler with the de-skew branch (SR:364-434)."""

    def __init__(self):
        import shutil
        import tempfile
        self._dir = tempfile.mkdtemp(prefix="refsr_")
        so = os.path.join(self._dir, "libref_sr_imu.so")
        shutil.copy(os.path.join(_DIR, "libref_sr.so"), so)
        L = C.CDLL(so)
        vp = C.c_void_p
        L.ref_sr_process.argtypes = [vp, C.c_int, C.c_double]
        L.ref_sr_cloud.argtypes = [C.c_int, vp, C.c_int]
        L.ref_sr_imu.argtypes = [C.c_double, vp, vp, vp]
        L.ref_sr_imu_trans.argtypes = [vp]
        self.L = L

    def imu(self, stamp, quat_xyzw, angular_velocity, linear_acceleration):
        q, a, l = (np.ascontiguousarray(v, np.float64) for v in (quat_xyzw, angular_velocity, linear_acceleration))
        self.L.ref_sr_imu(float(stamp), q.ctypes.data, a.ctypes.data, l.ctypes.data)

    def process(self, xyz, stamp):
        xyz = np.ascontiguousarray(xyz, np.float32)
        self.L.ref_sr_process(xyz.ctypes.data, xyz.shape[0], float(stamp))
        tr = np.zeros(12, np.float32)
        self.L.ref_sr_imu_trans(tr.ctypes.data)
        return [_cloud(self.L.ref_sr_cloud, w) for w in range(5)], tr


class LoWithImu:
    """A PRIVATE copy of the reference's laserOdometry.cpp (its state lives in file-scope globals) stepped with the twelve
    floats of /imu_trans: the IMU formulas of LO:201-225, 385-409, 566-568, 1053-1064."""

    def __init__(self):
        import shutil
        import tempfile
        self._dir = tempfile.mkdtemp(prefix="reflo_")
        so = os.path.join(self._dir, "libref_lo_imu.so")
        shutil.copy(os.path.join(_DIR, "libref_lo.so"), so)
        L = C.CDLL(so)
        vp = C.c_void_p
        L.ref_lo_step.argtypes = [vp, C.c_int, vp, C.c_int, vp, C.c_int, vp, C.c_int, vp, C.c_int, vp, C.c_double, vp]
        L.ref_lo_cloud.argtypes = [C.c_int, vp, C.c_int]
        L.ref_lo_start()
        self.L = L

    def step(self, feat, stamp, imu12):
        full, sharp, less_sharp, flat, less_flat = (np.ascontiguousarray(x, np.float32) for x in feat)
        imu = np.ascontiguousarray(imu12, np.float32)
        out = np.zeros(18, np.float32)
        self.L.ref_lo_step(sharp.ctypes.data, sharp.shape[0], less_sharp.ctypes.data, less_sharp.shape[0], flat.ctypes.data, flat.shape[0],
                           less_flat.ctypes.data, less_flat.shape[0], full.ctypes.data, full.shape[0], imu.ctypes.data, float(stamp),
                           out.ctypes.data)
        clouds = [_cloud(self.L.ref_lo_cloud, w) for w in range(3)] if out[13] > 0 else None
        return out, clouds

    def close(self):
        self.L.ref_lo_stop()


class PrivateNodes:
    """PRIVATE copies of all three reference nodes (their state -- incl. scanRegistration's never re-initialised static
    arrays -- lives in file-scope globals and would otherwise depend on what ran before in this process), wired like
    `process()`: one sweep through scanRegistration -> laserOdometry -> laserMapping."""

    def __init__(self):
        import shutil
        import tempfile
        self._dir = tempfile.mkdtemp(prefix="refnodes_")
        libs = []
        for name in ("sr", "lo", "lm"):
            so = os.path.join(self._dir, f"libref_{name}_private.so")
            shutil.copy(os.path.join(_DIR, f"libref_{name}.so"), so)
            libs.append(C.CDLL(so))
        sr, lo, lm = libs
        vp, ip = C.c_void_p, C.POINTER(C.c_int)
        sr.ref_sr_process.argtypes = [vp, C.c_int, C.c_double]
        sr.ref_sr_cloud.argtypes = [C.c_int, vp, C.c_int]
        lo.ref_lo_step.argtypes = [vp, C.c_int, vp, C.c_int, vp, C.c_int, vp, C.c_int, vp, C.c_int, vp, C.c_double, vp]
        lo.ref_lo_odometry.argtypes = [vp]
        lo.ref_lo_cloud.argtypes = [C.c_int, vp, C.c_int]
        lo.ref_lo_control.argtypes = [C.c_int]
        lm.ref_lm_odometry_only.argtypes = [vp, C.c_double]
        lm.ref_lm_step.argtypes = [vp, C.c_int, vp, C.c_int, vp, C.c_int, vp, C.c_double, vp]
        lm.ref_lm_map_size.argtypes = [ip, ip]
        lm.ref_lm_cloud.argtypes = [C.c_int, vp, C.c_int]
        lo.ref_lo_start()
        lm.ref_lm_start()
        self.sr, self.lo, self.lm = sr, lo, lm

    def lm_cloud(self, which):
        """The last published /laser_cloud_surround (0) or /velodyne_cloud_registered (1)."""
        return _cloud(self.lm.ref_lm_cloud, which)

    def control_reset(self):
        """IMControl{systemInited=false} to laserOdometry (IN:281-284)."""
        self.lo.ref_lo_control(0)

    def process(self, xyz, stamp):
        r = SweepOut()
        xyz = np.ascontiguousarray(xyz, np.float32)
        self.sr.ref_sr_process(xyz.ctypes.data, xyz.shape[0], float(stamp))
        full, sharp, less_sharp, flat, less_flat = feat = [_cloud(self.sr.ref_sr_cloud, w) for w in range(5)]
        r.features = feat
        o = np.zeros(18, np.float32)
        self.lo.ref_lo_step(sharp.ctypes.data, sharp.shape[0], less_sharp.ctypes.data, less_sharp.shape[0], flat.ctypes.data, flat.shape[0],
                            less_flat.ctypes.data, less_flat.shape[0], full.ctypes.data, full.shape[0], None, float(stamp), o.ctypes.data)
        pose7 = np.zeros(7, np.float64)
        self.lo.ref_lo_odometry(pose7.ctypes.data)
        r.odom, r.rel = o[:6].copy(), o[6:12].copy()
        r.odom_published, r.clouds_published, r.fullres_published = bool(o[12] > 0), bool(o[13] > 0), bool(o[14] > 0)
        r.mapping_ran, r.mapped, r.bef_mapped, r.tobe_mapped = False, None, None, None
        if r.odom_published:
            if not r.fullres_published:
                self.lm.ref_lm_odometry_only(pose7.ctypes.data, float(stamp))
            else:
                c, s, f = [_cloud(self.lo.ref_lo_cloud, w) for w in range(3)]
                m = np.zeros(24, np.float32)
                self.lm.ref_lm_step(c.ctypes.data, c.shape[0], s.ctypes.data, s.shape[0], f.ctypes.data, f.shape[0], pose7.ctypes.data,
                                    float(stamp), m.ctypes.data)
                r.mapping_ran = True
                r.mapped, r.bef_mapped, r.tobe_mapped = m[:6].copy(), m[6:12].copy(), m[12:18].copy()
        return r

    def map_size(self):
        a, b = C.c_int(), C.c_int()
        self.lm.ref_lm_map_size(C.byref(a), C.byref(b))
        return a.value, b.value

    def close(self):
        self.lo.ref_lo_stop()
        self.lm.ref_lm_stop()
