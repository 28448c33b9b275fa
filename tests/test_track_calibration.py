"""track calibration (weighted Kabsch + re-weighting loop + pairwise smoothing; TC, WC, LD:57-83).

CPU: the oracle restatement (oracle/orc_track.py) against the reference's own track_calibration.cc / weight_calculation.cc
(oracle/_ref/libref_tc.so, compiled against a minimal MatrixXd shim) bit for bit, the restated Jacobi SVD against LAPACK,
and the product's host arithmetic (loam_track_* inside libloamgpu.so) against the oracle bit for bit; the O(N) closed
form of the smoothing against the exact loop within 1e-9 m.  GPU: the smoothing kernel (every point's sum in the serial
loop's order) equal to the oracle bit for bit, alone and through the whole LD:57-83 loop.
Tolerances: fp64 track arithmetic is compared with == everywhere except the closed form (1e-9 m absolute at UTM-sized
coordinates ~ 3e-16 relative) and the LAPACK cross-check (1e-12)."""
import numpy as np
import pytest


def make_tracks(n, seed, noise=1.5, outliers=0):
    """A driving-like SLAM track (local frame) and the ENU track it should align to (rotated, UTM-sized offsets, GPS noise)."""
    rng = np.random.default_rng(seed)
    t = 1.5e9 + np.arange(n) * 1.0
    speed = 1.5 + 0.8 * np.sin(0.05 * np.arange(n))  # below and above SPEED = 2.2 m per sample
    head = np.cumsum(0.01 + 0.004 * rng.standard_normal(n))
    x, y = np.cumsum(speed * np.cos(head)), np.cumsum(speed * np.sin(head))
    slam = np.stack([x, y, np.full(n, 10.0), t], 1)
    th = 0.7
    R = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    e = (R @ slam[:, :2].T).T + np.array([5.0e5, 3.2e6]) + rng.standard_normal((n, 2)) * noise
    for k in rng.choice(n, size=min(outliers, n), replace=False):
        e[k] += rng.standard_normal(2) * 25.0  # multipath-like jumps the re-weighting has to suppress
    enu = np.stack([e[:, 0], e[:, 1], 40.0 + rng.standard_normal(n) * 3.0, t], 1)
    return slam, enu


SIZES = [(1, 0), (2, 1), (3, 2), (17, 3), (300, 4), (1201, 5)]


def _need_ref():
    from oracle import ref
    if not ref.tc_available():
        pytest.skip("oracle/_ref/libref_tc.so not built (needs /root/reference)")
    return ref


@pytest.mark.parametrize("n,seed", SIZES)
def test_oracle_matches_reference_track_calibration(n, seed):
    ref = _need_ref()
    from oracle import orc_track
    slam, enu = make_tracks(n, seed, outliers=n // 40)
    w_ref = ref.tc_speed_weights(slam)
    w = orc_track.speed_weights(slam)
    assert np.array_equal(w, w_ref)
    cal_ref, rot_ref = ref.tc_calibrate(slam, enu, w_ref)
    tc = orc_track.TrackCalibration(slam, enu, w)
    tc.do_icp()
    assert np.array_equal(tc.rotated, rot_ref)
    cal = tc.do_calibration()
    assert np.array_equal(cal, cal_ref)
    w2_ref = ref.tc_residual_weights(slam, enu, cal_ref)
    assert np.array_equal(orc_track.residual_weights(slam, enu, cal), w2_ref)


def test_oracle_long_loop_matches_reference():
    ref = _need_ref()
    from oracle import orc_track
    slam, enu = make_tracks(400, 11, outliers=12)
    w_ref, cal_ref = ref.tc_long(slam, enu, 5)
    w, cal = orc_track.calibrate_long(slam, enu, 5)
    assert np.array_equal(w, w_ref) and np.array_equal(cal, cal_ref)


def test_restated_svd_against_lapack(orc):
    rng = np.random.default_rng(0)
    for k in range(200):
        H = rng.standard_normal((3, 3)) * 10.0 ** rng.integers(-3, 6)
        if k % 4 == 0:  # the planar case of the track problem: third row and column exactly zero
            H[2, :] = 0.0
            H[:, 2] = 0.0
        U, S, V = orc.svd3(H)
        scale = np.abs(H).max()
        assert np.abs(U @ np.diag(S) @ V.T - H).max() <= 1e-13 * scale
        assert np.abs(U.T @ U - np.eye(3)).max() <= 1e-13 and np.abs(V.T @ V - np.eye(3)).max() <= 1e-13
        assert np.abs(S - np.linalg.svd(H, compute_uv=False)).max() <= 1e-12 * scale
        assert S[0] >= S[1] >= S[2] >= 0.0
        if k % 4 == 0:
            assert U[2, 2] == 1.0 and V[2, 2] == 1.0


def test_alignment_against_lapack_kabsch():
    """The weighted rigid fit recovers the generating rotation; equal to the textbook Kabsch solution via LAPACK."""
    from oracle import orc_track
    slam, enu = make_tracks(500, 21, noise=0.5)
    w = orc_track.speed_weights(slam)
    tc = orc_track.TrackCalibration(slam, enu, w)
    T = tc.do_icp()
    A, B = tc.slam[:, :2], tc.enu[:, :2]
    ca, cb = (A * w[:, None]).sum(0) / w.sum(), (B * w[:, None]).sum(0) / w.sum()
    H = ((A - ca) * (w ** 2)[:, None]).T @ (B - cb)
    U, _, Vt = np.linalg.svd(H)
    Rk = Vt.T @ U.T
    assert np.linalg.det(Rk) > 0
    assert np.abs(Rk - T[:2, :2]).max() <= 1e-12
    assert abs(np.arctan2(T[1, 0], T[0, 0]) - 0.7) < 5e-3


@pytest.mark.parametrize("n,seed", SIZES)
def test_product_host_arithmetic_equals_oracle(n, seed):
    from gpscalibration_b200 import capi
    from oracle import orc, orc_track
    slam, enu = make_tracks(n, seed, outliers=n // 40)
    w = capi.track_speed_weights(slam)
    assert np.array_equal(w, orc_track.speed_weights(slam))
    tc = capi.TrackCalibration(slam, enu, w, mode=1)
    tc.do_icp()
    o = orc_track.TrackCalibration(slam, enu, w)
    o.do_icp()
    assert np.array_equal(tc.T, o.T) and np.array_equal(tc.rotated, o.rotated)
    exact = o.do_calibration()
    closed = tc.do_calibration()  # mode 1: O(N) closed form on the host
    assert np.abs(closed - exact).max() <= 1e-9
    assert np.array_equal(capi.track_residual_weights(slam, enu, exact), orc_track.residual_weights(slam, enu, exact))
    H = np.random.default_rng(seed).standard_normal((3, 3))
    for a, b in zip(capi.track_svd3(H), orc.svd3(H)):
        assert np.array_equal(a, b)


def test_product_rejects_mismatched_tracks():
    from gpscalibration_b200 import capi
    slam, enu = make_tracks(10, 1)
    with pytest.raises(ValueError):
        capi.TrackCalibration(slam, enu[:9], np.ones(10))
    assert capi.load_library().loam_track_speed_weights(None, 3, None) == -1


def test_long_loop_closed_form_close_to_oracle():
    from gpscalibration_b200 import capi
    from oracle import orc_track
    slam, enu = make_tracks(300, 31, outliers=8)
    w, cal = capi.track_calibrate_long(slam, enu, 5, mode=1)
    w0, cal0 = orc_track.calibrate_long(slam, enu, 5)
    assert np.abs(cal - cal0).max() <= 1e-7 and np.abs(w - w0).max() <= 1e-6 * np.abs(w0).max()
    # the re-weighting does its job: the final weights of the injected outliers are the smallest
    resid = np.hypot(*(enu[:, :2] - cal0[:, :2]).T)
    worst = np.argsort(resid[1:-1])[-8:] + 1  # (first weight is 1 by definition, the last 0 by the clamp fence)
    assert w0[worst].max() < np.median(w0[1:-1])


@pytest.mark.gpu
@pytest.mark.parametrize("n,seed", SIZES + [(5000, 7)])
def test_gpu_smoothing_kernel_equals_oracle(n, seed):
    from gpscalibration_b200 import capi
    from oracle import orc_track
    slam, enu = make_tracks(n, seed, outliers=n // 40)
    w = capi.track_speed_weights(slam)
    tc = capi.TrackCalibration(slam, enu, w, mode=0)
    tc.do_icp()
    cal = tc.do_calibration()
    o = orc_track.TrackCalibration(slam, enu, w)
    o.do_icp()
    assert np.array_equal(cal, o.do_calibration())


@pytest.mark.gpu
def test_gpu_long_loop_equals_oracle_and_reference():
    from gpscalibration_b200 import capi
    from oracle import orc_track, ref
    slam, enu = make_tracks(600, 41, outliers=15)
    w, cal = capi.track_calibrate_long(slam, enu, 5, mode=0)
    w0, cal0 = orc_track.calibrate_long(slam, enu, 5)
    assert np.array_equal(w, w0) and np.array_equal(cal, cal0)
    gold = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "ref_track_calibration.npz"))
    w2, cal2 = capi.track_calibrate_long(gold["slam"], gold["enu"], 5, mode=0)
    assert np.array_equal(w2, gold["w"]) and np.array_equal(cal2, gold["cal"])  # written by the reference's own code


def test_oracle_reproduces_reference_golden_track():
    import os
    from oracle import orc_track
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_track_calibration.npz"))
    w, cal = orc_track.calibrate_long(gold["slam"], gold["enu"], 5)
    assert np.array_equal(w, gold["w"]) and np.array_equal(cal, gold["cal"])
