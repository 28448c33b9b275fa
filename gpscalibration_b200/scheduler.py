"""Segment scheduler: host-side mirror of input_data.cpp's replay loop on top of loam_replay_segments.

The reference's node replays a list of bags twice -- long tracks, then short overlapping tracks -- measuring the
travelled distance on the SLAM output, resetting the SLAM nodes at every cut and publishing the finished tracks on
/slam_track (IN:244-446).  `SegmentScheduler` does the same against any object with `control()` and
`publish(bag, msg) -> (stamp, odometry or None)`; `GpuSlam` is that object for the CUDA pipeline.  `run(parallel=True)`
gives each pass its own pipeline (its own GPU when two devices are passed); that mode is NOT bit-identical to the
reference: there the last pose of pass 0 leaks into pass 1 (preOdometry is only cleared inside the replay loop,
IN:362) and cuts a one-pose track first, which a pass started on its own does not produce.
"""
import ctypes as C
import threading

import numpy as np

from . import capi

_PUBLISH = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int))
_CONTROL = C.CFUNCTYPE(None, C.c_void_p)
_TRACK = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.POINTER(C.c_double), C.c_int)


class _Callbacks(C.Structure):
    _fields_ = [("publish", _PUBLISH), ("control", _CONTROL), ("slam_track", _TRACK), ("user", C.c_void_p)]


class ReplayStats(C.Structure):
    _fields_ = [("published", C.c_longlong), ("lost", C.c_longlong), ("resets", C.c_longlong), ("tracks", C.c_longlong)]


def replay_segments(messages_per_bag, long_distance, short_distance, overlap_distance, slam, passes=(0, 1)):
    """One call of loam_replay_segments; returns ([(track_flag, ndarray (n, 4) x y z t), ...], stats)."""
    lib = capi.load_library()
    lib.loam_replay_segments.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.POINTER(_Callbacks),
                                         C.POINTER(ReplayStats)]
    tracks, errors = [], []

    def publish(user, bag, msg, stamp, odo, arrived):
        try:
            s, o = slam.publish(bag, msg)
            stamp[0] = s
            arrived[0] = 0 if o is None else 1
            if o is not None:
                for i in range(4):
                    odo[i] = o[i]
        except BaseException as e:  # never let an exception cross the C frame
            errors.append(e)
            arrived[0] = 0

    def control(user):
        try:
            slam.control()
        except BaseException as e:
            errors.append(e)

    def slam_track(user, flag, xyzt, n):
        tracks.append((flag, np.ctypeslib.as_array(xyzt, shape=(n, 4)).copy() if n else np.zeros((0, 4))))

    cb = _Callbacks(_PUBLISH(publish), _CONTROL(control), _TRACK(slam_track), None)
    counts = (C.c_int * len(messages_per_bag))(*messages_per_bag)
    stats = ReplayStats()
    rc = lib.loam_replay_segments(counts, len(messages_per_bag), long_distance, short_distance, overlap_distance, min(passes), max(passes),
                                  C.byref(cb), C.byref(stats))
    if errors:
        raise errors[0]
    if rc:
        raise capi.LoamError(rc, "loam_replay_segments")
    return tracks, stats


class GpuSlam:
    """scanRegistration + laserOdometry + laserMapping + transformMaintenance on one handle, fed from in-memory bags."""

    def __init__(self, bags, stamps, device=0, **params):
        self.bags, self.stamps = bags, stamps
        self.gpu = capi.LoamGpu(device=device, **params)
        self.sweeps = 0

    def control(self):  # IMControl{systemInited=false}: LO:411-415 (LM and TM re-initialise on the zero pose, LM:316-319 / TM:264-275)
        self.gpu.reset()

    def publish(self, bag, msg):
        stamp = self.stamps[bag][msg]
        r = self.gpu.process_sweep(self.bags[bag][msg], stamp)
        self.sweeps += 1
        if not r.odom.odom_published:
            return stamp, None
        _, track = self.gpu.integrate_odometry(r.odom.transform_sum, stamp)  # /true_odometry_to_init (TM:262-315)
        if r.mapping_ran:
            self.gpu.integrate_mapping(r.map.transform_aft_mapped, r.map.transform_bef_mapped)  # TM:317-338
        return stamp, tuple(track)


class SegmentScheduler:
    def __init__(self, long_distance, short_distance, overlap_distance):
        self.distances = (float(long_distance), float(short_distance), float(overlap_distance))

    def run(self, bags, stamps, devices=(0,), parallel=False, slam_factory=None, **params):
        """bags: list of lists of (n, 3) sweeps; stamps: matching header stamps.  Returns the /slam_track messages in
        the reference's order (all of pass 0, then all of pass 1) and the per-pass statistics."""
        counts = [len(b) for b in bags]
        make = slam_factory or (lambda dev: GpuSlam(bags, stamps, device=dev, **params))
        if not parallel:
            slam = make(devices[0])
            tracks, stats = replay_segments(counts, *self.distances, slam, passes=(0, 1))
            return tracks, [stats]
        out = [None, None]

        def one(p):
            slam = make(devices[p % len(devices)])
            out[p] = replay_segments(counts, *self.distances, slam, passes=(p, p))

        th = [threading.Thread(target=one, args=(p,)) for p in (0, 1)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        if out[0] is None or out[1] is None:
            raise RuntimeError("a replay pass failed")
        return out[0][0] + out[1][0], [out[0][1], out[1][1]]
