"""The IMU branch of scanRegistration (imuHandler SR:754-837, AccumulateIMUShift SR:187-233, the per-point de-skew SR:364-434
with SR:121-184; dormant in the shipped pipeline but live code).  CPU: the oracle restatement against the reference's own
scanRegistration.cpp (a private copy of oracle/_ref/libref_sr.so: the node keeps its IMU state in globals) bit for bit --
clouds and the 12 floats of /imu_trans -- on IMU streams that exercise both interpolation branches, the 0.2 s staleness
gates, the yaw filter, the acceleration gate and a yaw wrap.  GPU: loam_imu_push + loam_extract against the oracle."""
import numpy as np
import pytest


def quat(roll, pitch, yaw):
    cy, sy, cp, sp, cr, sr = np.cos(yaw / 2), np.sin(yaw / 2), np.cos(pitch / 2), np.sin(pitch / 2), np.cos(roll / 2), np.sin(roll / 2)
    return np.array([sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy, cr * cp * cy + sr * sp * sy])


def imu_stream(t0, t1, rate, seed, yaw0=0.0, gap=None):
    """(stamp, quaternion, angular velocity, linear acceleration) at `rate` Hz; `gap` = (a, b): no messages in that interval."""
    rng = np.random.default_rng(seed)
    out = []
    t = t0
    while t < t1:
        if not (gap and gap[0] <= t < gap[1]):
            yaw = yaw0 + 0.35 * (t - t0) + 0.02 * np.sin(3 * t)
            yaw = (yaw + np.pi) % (2 * np.pi) - np.pi  # wraps through +-pi on long streams
            roll, pitch = 0.03 * np.sin(1.3 * t), 0.02 * np.cos(0.7 * t)
            wz = 0.35 + 0.06 * np.cos(3 * t) if rng.random() > 0.1 else rng.choice([4.0, -4.0, 0.0])  # the yaw filter's branches
            la = np.array([0.4 * np.sin(t) - np.sin(pitch) * 9.81, 0.3 * np.cos(2 * t) + np.sin(roll) * np.cos(pitch) * 9.81,
                           np.cos(roll) * np.cos(pitch) * 9.81 + 0.1 * rng.standard_normal()])
            if rng.random() < 0.03:
                la[0] += 5.0  # trips the |acc| > 2 gate: the message's integration is skipped
            out.append((t, quat(roll, pitch, yaw), np.array([0.0, 0.0, wz]), la))
        t += 1.0 / rate
    return out


def scenario(seed, sensor=0):
    """Sweeps at 10 Hz with stamps, IMU messages delivered before each sweep up to a little PAST its end (so both the
    'IMU newer' and the 'IMU older' branch run), one gap longer than 0.2 s, one sweep whose first point is invalid."""
    from gpscalibration_b200 import SweepGenerator
    gen = SweepGenerator(sensor=sensor, scene=seed % 2, seed=0xC0FFEE + seed)
    msgs = imu_stream(99.95, 101.3, 100.0, seed, yaw0=2.9, gap=(100.52, 100.78))
    events = []
    k_msg = 0
    for k in range(10):
        stamp = 100.0 + 0.1 * k
        lead = 0.13 if k % 3 else 0.04  # how far past the sweep's stamp the IMU has been heard
        while k_msg < len(msgs) and msgs[k_msg][0] <= stamp + lead:
            events.append(("imu",) + msgs[k_msg])
            k_msg += 1
        xyz = gen.sweep(k)[0].copy()
        if k == 4:
            xyz[0] = np.nan  # the first finite point moves
        if k == 6:
            xyz[0, 2] = 40.0  # the first point leaves the ring table: the Start values are NOT latched in this sweep
        events.append(("sweep", stamp, xyz))
    return events


@pytest.mark.parametrize("seed", [0, 1])
def test_oracle_imu_branch_matches_reference(orc, seed):
    from oracle import ref
    import os
    if not os.path.exists(os.path.join(ref._DIR, "libref_sr.so")):
        pytest.skip("oracle/_ref/libref_sr.so not built (needs /root/reference)")
    r = ref.SrWithImu()
    o = orc.ScanRegistration()
    n_sweeps = 0
    for ev in scenario(seed):
        if ev[0] == "imu":
            r.imu(*ev[1:])
            o.imu(*ev[1:])
        else:
            _, stamp, xyz = ev
            rc, rtr = r.process(xyz, stamp)
            oc, otr = o.extract_imu(xyz, stamp)
            assert np.array_equal(rtr.view(np.uint32), otr.view(np.uint32)), (n_sweeps, rtr, otr)
            for name, a in zip(o.CLOUDS, rc):
                b = oc[name]
                assert a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32)), (n_sweeps, name)
            n_sweeps += 1
    assert n_sweeps == 10 and np.abs(otr).max() > 1e-3  # the de-skew did something


@pytest.mark.parametrize("seed", [0, 1])
def test_oracle_reproduces_reference_imu_golden(orc, seed):
    """tests/golden/ref_imu_seq.npz was written by the reference's own scanRegistration.cpp (IMU branch) and laserOdometry.cpp
    (tests/golden/make_golden_imu.py); it travels to boxes without /root/reference.  The oracle reproduces it exactly:
    /imu_trans, sizes and hashes of the five clouds, the odometry node's poses, flags and published clouds."""
    import hashlib
    import os

    def sha(a):
        return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_imu_seq.npz"))
    p = f"s{seed}_"
    osr, olo = orc.ScanRegistration(), orc.LaserOdometry()
    in_hash, k = hashlib.sha256(), 0
    for ev in scenario(seed):
        if ev[0] == "imu":
            for v in ev[1:]:
                in_hash.update(np.ascontiguousarray(v, np.float64).tobytes())
            osr.imu(*ev[1:])
            continue
        _, stamp, xyz = ev
        in_hash.update(np.ascontiguousarray(xyz, np.float32).tobytes())
        oc, otr = osr.extract_imu(xyz, stamp)
        feat = [oc[n] for n in ("full", "sharp", "less_sharp", "flat", "less_flat")]
        assert np.array_equal(otr.view(np.uint32), g[p + "imu_trans"][k].view(np.uint32)), k
        assert [f.shape[0] for f in feat] == g[p + "counts"][k].tolist(), k
        assert [sha(f) for f in feat] == [str(h) for h in g[p + "cloud_hash"][k]], k
        out, clouds = olo.step(feat, otr)
        assert np.array_equal(out[:15].view(np.uint32), g[p + "lo_out"][k].view(np.uint32)), k
        want_hash = [str(h) for h in g[p + "lo_cloud_hash"][k]]
        if clouds is None:
            assert want_hash == ["", "", ""], k
        else:
            # /velodyne_cloud_3 is only republished every second sweep: the reference's capture keeps the last one
            n_cmp = 3 if out[14] > 0 else 2
            assert [sha(c) for c in clouds[:n_cmp]] == want_hash[:n_cmp], k
        k += 1
    assert k == 10 and in_hash.hexdigest() == str(g[p + "in_hash"])  # the scenario is still the one the fixture was made from


@pytest.mark.parametrize("seed", [0, 1])
def test_oracle_odometry_with_imu_trans_matches_reference(orc, seed):
    """The odometry node fed with a NON-ZERO /imu_trans (LO:201-225 TransformToEnd's IMU rotations, LO:385-409 the handler,
    LO:566-568 the velocity prior, LO:1053-1064 the shift and PluginIMURotation): the oracle's restatement against the
    reference's own laserOdometry.cpp (a private copy of libref_lo.so), bit for bit on transformSum, the sweep-relative
    transform, the publish flags and the three published clouds.  Features and the twelve floats come from the reference's
    own scanRegistration.cpp running its IMU branch."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    sr, lo = ref.SrWithImu(), ref.LoWithImu()
    o = orc.LaserOdometry()
    try:
        n_sweeps, moved, imu_seen = 0, False, 0.0
        for ev in scenario(seed):
            if ev[0] == "imu":
                sr.imu(*ev[1:])
                continue
            _, stamp, xyz = ev
            feat, tr = sr.process(xyz, stamp)
            want, wclouds = lo.step(feat, stamp, tr)
            got, gclouds = o.step(feat, tr)
            assert np.array_equal(want[:15].view(np.uint32), got[:15].view(np.uint32)), (n_sweeps, want, got)
            assert (wclouds is None) == (gclouds is None), n_sweeps
            if wclouds is not None:
                for w in range(3 if want[14] > 0 else 2):
                    assert wclouds[w].shape == gclouds[w].shape, (n_sweeps, w)
                    assert np.array_equal(wclouds[w].view(np.uint32), gclouds[w].view(np.uint32)), (n_sweeps, w)
            moved = moved or np.abs(want[:6]).max() > 1e-3
            imu_seen = max(imu_seen, float(np.abs(tr).max()))
            n_sweeps += 1
        assert n_sweeps == 10 and moved and imu_seen > 1e-3
    finally:
        lo.close()


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [0, 1])
def test_gpu_imu_deskew_equals_oracle(orc, seed):
    """loam_imu_push (host: imuHandler + AccumulateIMUShift) + loam_extract with the de-skew kernel (prefix-max scans in
    place of the reference's point-to-point state) against the oracle: all five clouds and /imu_trans bit for bit."""
    from gpscalibration_b200 import LoamGpu
    gpu = LoamGpu()
    o = orc.ScanRegistration()
    n_sweeps = 0
    for ev in scenario(seed):
        if ev[0] == "imu":
            gpu.imu_push(*ev[1:])
            o.imu(*ev[1:])
        else:
            _, stamp, xyz = ev
            c = gpu.extract(xyz, stamp)
            oc, otr = o.extract_imu(xyz, stamp)
            assert (c.n_full, c.n_sharp, c.n_less_sharp, c.n_flat, c.n_less_flat) == tuple(oc[k].shape[0] for k in o.CLOUDS), n_sweeps
            gtr = gpu.imu_trans()
            assert np.array_equal(gtr.view(np.uint32), otr.view(np.uint32)), (n_sweeps, gtr, otr)
            for name in o.CLOUDS:
                a, b = gpu.cloud(name), oc[name]
                assert a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32)), (n_sweeps, name)
            n_sweeps += 1
    assert n_sweeps == 10
    gpu.close()


@pytest.mark.gpu
def test_gpu_imu_deskew_with_empty_rings(orc):
    """The IMU branch on true VLP-16 angles: rings 6, 8 and 10 of the reference's table stay empty, so every sweep also goes
    through the virtual-ring replay (SR:480-490) -- on de-skewed points.  Clouds and /imu_trans equal to the oracle."""
    from gpscalibration_b200 import LoamGpu
    gpu = LoamGpu()
    o = orc.ScanRegistration()
    n_sweeps = 0
    for ev in scenario(2, sensor=1):
        if ev[0] == "imu":
            gpu.imu_push(*ev[1:])
            o.imu(*ev[1:])
        else:
            _, stamp, xyz = ev
            c = gpu.extract(xyz, stamp)
            oc, otr = o.extract_imu(xyz, stamp)
            assert (c.n_full, c.n_sharp, c.n_less_sharp, c.n_flat, c.n_less_flat) == tuple(oc[k].shape[0] for k in o.CLOUDS), n_sweeps
            assert np.array_equal(gpu.imu_trans().view(np.uint32), otr.view(np.uint32)), n_sweeps
            for name in o.CLOUDS:
                a, b = gpu.cloud(name), oc[name]
                assert a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32)), (n_sweeps, name)
            n_sweeps += 1
    assert n_sweeps == 10 and (o.ints("scan_start")[[6, 8, 10]] == 0).all()  # the three rings really were empty
    gpu.close()


@pytest.mark.gpu
@pytest.mark.skip(reason="written after this round's GPU minutes were spent: never run on a B200 yet -- run it first next round")
def test_gpu_odometry_with_imu_equals_oracle(orc):
    """Node level: loam_imu_push + loam_extract + loam_odometry_process against the oracle's scanRegistration (IMU branch) +
    odometry node fed with its /imu_trans -- transformSum, the sweep-relative transform and the flags, bit for bit.  (The
    oracle side of this comparison is pinned against the reference's own code by
    test_oracle_odometry_with_imu_trans_matches_reference; the kernels involved are covered by test_transform_to_end_parity
    with a non-zero imu_trans and by the extraction tests above.)"""
    from gpscalibration_b200 import LoamGpu
    gpu = LoamGpu()
    osr, olo = orc.ScanRegistration(), orc.LaserOdometry()
    for ev in scenario(1):
        if ev[0] == "imu":
            gpu.imu_push(*ev[1:])
            osr.imu(*ev[1:])
            continue
        _, stamp, xyz = ev
        gpu.extract(xyz, stamp)
        got = gpu.odometry_process()
        oc, otr = osr.extract_imu(xyz, stamp)
        want, _ = olo.step([oc[k] for k in ("full", "sharp", "less_sharp", "flat", "less_flat")], otr)
        assert np.array_equal(np.array(got.transform_sum, np.float32).view(np.uint32), want[:6].view(np.uint32))
        assert np.array_equal(np.array(got.transformation, np.float32).view(np.uint32), want[6:12].view(np.uint32))
        assert (got.odom_published, got.clouds_published, got.fullres_published) == tuple(int(v) for v in want[12:15])
    gpu.close()


@pytest.mark.gpu
def test_gpu_pipeline_runs_with_imu():
    """The whole path with IMU messages: the de-skewed clouds and /imu_trans feed odometry and mapping of the same handle
    (LO:201-225, 566-568, 1053-1064 use the 12 floats); without messages nothing changes."""
    from gpscalibration_b200 import LoamGpu
    with_imu, without = LoamGpu(), LoamGpu()
    last = None
    for ev in scenario(0):
        if ev[0] == "imu":
            with_imu.imu_push(*ev[1:])
        else:
            _, stamp, xyz = ev
            a = with_imu.process_sweep(xyz, stamp)
            b = without.process_sweep(xyz, stamp)
            last = (a, b)
    a, b = last
    assert a.odom.odom_published and b.odom.odom_published
    assert np.isfinite(np.array(a.odom.transform_sum)).all()
    assert list(a.odom.transform_sum) != list(b.odom.transform_sum)  # the IMU prior and the de-skew changed the registration
    assert np.abs(with_imu.imu_trans()).max() > 1e-3 and np.abs(without.imu_trans()).max() == 0.0
    with_imu.close()
    without.close()


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["submit", "batch", "lockstep"])
def test_gpu_pipelined_imu_equals_blocking(mode):
    """loam_pipeline_imu_push: the messages are applied by the extraction stage in submission order and /imu_trans travels
    with the features to the odometry stage -- every result field of the pipelined modes equals loam_imu_push +
    loam_process_sweep on one handle."""
    from gpscalibration_b200 import LoamGpu, LoamGpuPipeline
    from gpscalibration_b200.capi import pipeline_submit_batch
    blocking, pipe = LoamGpu(want_registered=1, want_surround=1), LoamGpuPipeline(want_registered=1, want_surround=1)
    want, n = [], 0
    for ev in scenario(1):
        if ev[0] == "imu":
            blocking.imu_push(*ev[1:])
            pipe.imu_push(*ev[1:])
        else:
            _, stamp, xyz = ev
            want.append(blocking.process_sweep(xyz, stamp))
            if mode == "submit":
                pipe.submit(xyz, stamp)
            else:
                pipeline_submit_batch([pipe], [xyz], lockstep=(mode == "lockstep"), stamps=[stamp])
            n += 1
    got = [pipe.wait() for _ in range(n)]
    moved = False
    for k, (a, b) in enumerate(zip(want, got)):
        assert bytes(a) == bytes(b), (mode, k)
        moved = moved or any(abs(v) > 1e-3 for v in a.odom.transform_sum)
    assert moved
    blocking.close()
    pipe.close()
