"""transformMaintenance (pose fusion + height-compensated track).  Host-only arithmetic:
the oracle restatement against the reference's own transformMaintenance.cpp (oracle/_ref), and the product's
loam_integrate_* against the oracle (needs a handle, hence GPU-marked)."""
import os

import numpy as np
import pytest


def _poses(n=40):
    rng = np.random.default_rng(3)
    T = np.zeros(6, np.float32)
    odo, maps = [], []
    for k in range(n):
        T = (T + np.array([0.001, 0.02, -0.0005, 0.05, -0.01, 1.0], np.float32) + rng.normal(0, 1e-3, 6).astype(np.float32)).astype(np.float32)
        odo.append(T.copy())
        aft = (T + rng.normal(0, 5e-3, 6)).astype(np.float32)
        maps.append((aft, T.copy()) if k % 2 == 1 else None)
    return odo, maps


def _same(a, b):
    return np.array_equal(a, b) or (np.isnan(a).any() and np.isnan(b).any())


def test_oracle_matches_reference_transform_maintenance(orc):
    from oracle import ref
    if not os.path.exists(os.path.join(ref._DIR, "libref_tm.so")):
        pytest.skip("oracle/_ref/libref_tm.so not built (needs /root/reference)")
    odo, maps = _poses()
    tm = orc.TransformMaintenance()
    zero = np.zeros(6, np.float32)
    L = orc.lib()

    def hop(t):  # the quaternion message hop is part of the reference node (TM:277-284)
        o = np.zeros(6, np.float32)
        a = np.ascontiguousarray(t, np.float32)
        L.orc_odometry_ros_hop(a.ctypes.data, o.ctypes.data)
        return o

    for seq in (0, 1):  # the second pass starts with a zero pose = reset (TM:264-275)
        for k, T in enumerate([zero] + odo):
            r_out, r_track = ref.tm_odometry(T, 10.0 + k * 0.1)
            o_out, o_track = tm.odometry(hop(T), 10.0 + k * 0.1)
            assert _same(r_out, o_out), (seq, k, r_out, o_out)
            assert _same(r_track, o_track), (seq, k, r_track, o_track)
            if k > 0 and maps[k - 1] is not None:
                aft, bef = maps[k - 1]
                ref.tm_aft_mapped(aft, bef, 10.0 + k * 0.1)
                tm.aft_mapped(hop(aft), bef)


@pytest.mark.gpu
def test_gpu_handle_integrate_equals_oracle(orc):
    from gpscalibration_b200 import LoamGpu
    gpu = LoamGpu()
    tm = orc.TransformMaintenance()
    odo, maps = _poses()
    for k, T in enumerate([np.zeros(6, np.float32)] + odo):
        a, ta = gpu.integrate_odometry(T, 5.0 + k)
        b, tb = tm.odometry(T, 5.0 + k)
        assert _same(a, b) and _same(ta, tb), k
        if k > 0 and maps[k - 1] is not None:
            gpu.integrate_mapping(*maps[k - 1])
            tm.aft_mapped(*maps[k - 1])
    gpu.close()
