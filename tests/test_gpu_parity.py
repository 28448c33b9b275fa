"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): integer / index results bit-exact (kNN tie rule: (d2, index) ascending);
fp32 normal equations within 1e-5 relative; poses within 1e-4 m / 1e-5 rad.
"""
import numpy as np
import pytest

from conftest import ulp_diff

pytestmark = pytest.mark.gpu


def test_exact_mode_is_on():
    """Exact mode is mandatory: the host libm the oracle calls (sinf cosf atanf atan2f) must be the glibc build the device
    ports restate, otherwise every bit-exact assert below would be meaningless.  Fails -- never loosens -- when it is not."""
    import tempfile
    from test_libm_port import run_check
    with tempfile.TemporaryDirectory() as d:
        assert run_check(d, 2) == [0, 0, 0, 0], "host libm differs from csrc/lg_libm.cuh: bit-exact parity cannot be checked on this box"
    print("exact_mode=1")


def _rand_cloud(rng, m, extent=30.0):
    p = np.empty((m, 4), np.float32)
    p[:, :3] = rng.uniform(-extent, extent, (m, 3))
    p[:, 3] = rng.uniform(0, 16, m)
    return p


# ------------------------------------------------------------------------------------------------ voxel grid (a5)
@pytest.mark.parametrize("m,leaf", [(0, 0.2), (1, 0.2), (37, 0.4), (2000, 0.2), (4096, 0.4), (4097, 0.2), (16384, 0.2),
                                    (20000, 0.4), (150000, 0.2)])
def test_voxel_grid_bit_exact(gpu, orc, m, leaf):
    rng = np.random.default_rng(m + 1)
    pts = _rand_cloud(rng, m, extent=8.0 if m < 50000 else 25.0)
    ref = orc.voxel_grid(pts, leaf)
    got = gpu.voxel_grid(pts, leaf)
    assert got.shape == ref.shape
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))  # centroids, order and count bit-exact


def test_voxel_grid_overflow_returns_input(gpu, orc):
    # > INT_MAX cells: PCL warns and hands the input back unfiltered
    pts = np.array([[0, 0, 0, 1], [3000, 3000, 3000, 2], [1, 1, 1, 3]], np.float32)
    ref = orc.voxel_grid(pts, 0.2)
    got = gpu.voxel_grid(pts, 0.2)
    assert np.array_equal(ref, pts) and np.array_equal(got, pts)


def test_voxel_grid_idempotent(gpu):
    rng = np.random.default_rng(7)
    once = gpu.voxel_grid(_rand_cloud(rng, 30000, 20.0), 0.4)
    twice = gpu.voxel_grid(once, 0.4)
    assert twice.shape[0] <= once.shape[0]
    thrice = gpu.voxel_grid(twice, 0.4)
    assert thrice.shape[0] <= twice.shape[0]


# ------------------------------------------------------------------------------------------------ extraction (a1-a4)
def _check_extract(gpu, orc_sr, xyz):
    ref = orc_sr.extract(xyz)
    c = gpu.extract(xyz)
    assert (c.n_full, c.n_sharp, c.n_less_sharp, c.n_flat, c.n_less_flat) == tuple(ref[k].shape[0] for k in orc_sr.CLOUDS)
    full = gpu.cloud("full")
    assert np.array_equal(full[:, :3], ref["full"][:, :3])  # ring-major order and coordinates bit-exact
    # intensity = ring + 0.1 * relTime goes through atan2f: the device port of glibc's atan2f is bit-identical
    # (tests/test_libm_port.py, test_exact_mode_is_on)
    assert np.array_equal(full[:, 3], ref["full"][:, 3])
    assert np.array_equal(gpu.diag("scan_start"), orc_sr.ints("scan_start"))
    assert np.array_equal(gpu.diag("scan_end"), orc_sr.ints("scan_end"))
    n = c.n_full
    assert np.array_equal(gpu.diag("curvature")[5:n - 5], orc_sr.curvature()[5:n - 5])
    assert np.array_equal(gpu.diag("picked_mask")[5:n - 5].astype(np.int32), orc_sr.ints("picked_mask")[5:n - 5])
    assert np.array_equal(gpu.diag("label")[5:n - 5].astype(np.int32), orc_sr.ints("label")[5:n - 5])
    for k in ("sharp", "less_sharp", "flat"):
        got = gpu.cloud(k)
        assert np.array_equal(got[:, :3], ref[k][:, :3]), k  # same points, same (ring, sector, pick) order
        assert np.array_equal(got[:, 3], ref[k][:, 3]), k
    lf = gpu.cloud("less_flat")
    assert np.array_equal(lf[:, :3], ref["less_flat"][:, :3])
    assert np.array_equal(lf[:, 3], ref["less_flat"][:, 3])
    return ref


def test_extract_vlp16_parity(gpu, orc, sweeps16):
    sr = orc.ScanRegistration()
    for k in (0, 1, 5):
        _check_extract(gpu, sr, sweeps16[k])


def test_extract_with_nan_and_ragged_input(gpu, orc, sweeps16):
    xyz = sweeps16[2].copy()
    rng = np.random.default_rng(3)
    bad = rng.choice(xyz.shape[0], 500, replace=False)
    xyz[bad[:250], 0] = np.nan
    xyz[bad[250:], 2] = np.inf
    xyz[0] = np.nan  # first point invalid: start azimuth comes from the first finite point
    xyz[-1] = np.nan
    xyz = xyz[:-777]  # ragged tail
    _check_extract(gpu, orc.ScanRegistration(), xyz)


def test_extract_hdl64_parity(orc):
    from gpscalibration_b200 import LoamGpu, SweepGenerator
    g = SweepGenerator(sensor=2, scene=0, seed=0xC0FFEE)
    step = 26.8 / 63.0
    gpu = LoamGpu(n_scans=64, ring_mode=1, ring_ang_min=-24.8, ring_ang_step=step)
    sr = orc.ScanRegistration(64, 1, -24.8, step)
    ref = _check_extract(gpu, sr, g.sweep(0)[0])
    assert ref["full"].shape[0] > 100000
    gpu.close()


def test_extract_dense_rings_large_sectors(gpu, orc):
    """Rings of 4200 points (sectors of ~700: the multi-chunk form of the selection walks) and curvature ties (a
    noise-free cylinder: equal curvatures -> the index decides, stable sort SR:568-576)."""
    rng = np.random.default_rng(11)
    per_ring, angles = 4200, np.array([-15, -13, -11, -9, -7, -5, -4, -3, -2, -1, 0, 1, 3, 5, 7, 9], np.float64)
    az = np.linspace(0.0, 2 * np.pi, per_ring, endpoint=False)
    cols = []
    for k in range(per_ring):  # column-major like a spinning sensor: all rings at one azimuth, then the next
        th = -az[k]
        for e in angles:
            rad = 8.0 + 2.5 * np.sin(3 * az[k]) + (0.6 if (k // 37) % 5 == 0 else 0.0)
            if e > 4:
                rad = 12.0  # noise-free cylinder: ties
            else:
                rad += rng.normal(0, 0.01)
            cols.append((rad * np.cos(th), rad * np.sin(th), rad * np.tan(np.deg2rad(e))))
    xyz = np.asarray(cols, np.float32)
    ref = _check_extract(gpu, orc.ScanRegistration(), xyz)
    assert ref["full"].shape[0] == xyz.shape[0] and ref["sharp"].shape[0] > 100


def test_extract_empty_and_tiny(gpu, orc):
    c = gpu.extract(np.zeros((0, 3), np.float32))
    assert c.n_full == 0 and c.n_sharp == 0 and c.n_less_flat == 0
    xyz = np.array([[5, 0, 0], [5, 1, 0], [5, 2, 0.1]], np.float32)
    ref = orc.ScanRegistration().extract(xyz)
    c = gpu.extract(xyz)
    assert c.n_full == ref["full"].shape[0] and c.n_sharp == 0 and c.n_flat == 0
    # 6..10 points: the curvature loop is empty but the last ring spans [0, n - 5) (SR:489-490) over never-initialised entries
    from gpscalibration_b200 import LoamGpu
    g2 = LoamGpu()
    sr = orc.ScanRegistration()
    for m in (8, 10, 11, 13):
        xyz = np.stack([np.full(m, 6.0), np.linspace(-0.5, 0.5, m), np.zeros(m)], 1).astype(np.float32)
        ref = sr.extract(xyz)
        c = g2.extract(xyz)
        assert (c.n_full, c.n_sharp, c.n_less_sharp, c.n_flat, c.n_less_flat) == tuple(ref[k].shape[0] for k in sr.CLOUDS), m
        for k in ("flat", "less_flat"):
            assert np.array_equal(g2.cloud(k), ref[k]), (m, k)
    g2.close()


def test_extract_true_vlp16_angles_empty_rings(gpu, orc):
    """True VLP-16 angles leave rings 6 / 8 / 10 of the reference's table empty and drop the +11/+13/+15 degree beams: the
    reference's scanStartInd / scanEndInd then overlap (SR:480-490): rings 5 / 7 / 9 are skipped and rings 6 / 8 / 10 span
    the cloud from index 0, re-sorting and re-picking over the rings before them, including the five leading entries of the
    static arrays that are never re-initialised (state carried from sweep to sweep).  Bit-exact, sweep after sweep."""
    from gpscalibration_b200 import SweepGenerator
    g = SweepGenerator(sensor=1, scene=0, seed=0xC0FFEE)
    sr = orc.ScanRegistration()
    for k in range(4):
        ref = _check_extract(gpu, sr, g.sweep(k)[0])
        assert np.array_equal(sr.ints("scan_start")[[6, 8, 10]], [0, 0, 0]) and (sr.ints("scan_end")[[5, 7, 9]] == 0).all()
        assert ref["less_flat"].shape[0] > ref["full"].shape[0] // 4  # the virtual rings contribute the earlier rings again


def test_extract_top_rings_missing(gpu, orc, sweeps16):
    """A sweep whose top three rings return nothing (sky): the last ring keeps scanStartInd 0 and spans the whole cloud."""
    sr = orc.ScanRegistration()
    for k in (3, 4):
        xyz = sweeps16[k]
        elev = np.degrees(np.arctan2(xyz[:, 2], np.hypot(xyz[:, 0], xyz[:, 1])))
        cut = xyz[elev < 4.0]
        assert cut.shape[0] < xyz.shape[0]
        _check_extract(gpu, sr, cut)
        assert sr.ints("scan_start")[15] == 0 and sr.ints("scan_end")[12] == 0
    _check_extract(gpu, sr, sweeps16[5])  # a fully populated sweep afterwards (the five stale entries keep their marks)


def test_extract_one_occluded_ring_mid_sequence(gpu, orc, sweeps16):
    sr = orc.ScanRegistration()
    _check_extract(gpu, sr, sweeps16[0])
    xyz = sweeps16[1]
    elev = np.degrees(np.arctan2(xyz[:, 2], np.hypot(xyz[:, 0], xyz[:, 1])))
    _check_extract(gpu, sr, xyz[np.abs(elev + 2.0) > 0.5])  # ring 8 (-2 degrees) occluded
    _check_extract(gpu, sr, xyz[elev > -14.0])               # ring 0 gone: every later ring keeps its own range
    _check_extract(gpu, sr, sweeps16[2])


def test_pipeline_parity_true_vlp16(orc):
    """extract -> odometry -> mapping on true VLP-16 angles (empty rings every sweep): poses, iteration counts and cloud
    sizes equal to the oracle bit for bit; the pipelined mode gives the same."""
    from gpscalibration_b200 import LoamGpu, LoamGpuPipeline, SweepGenerator
    g = SweepGenerator(sensor=1, scene=0, seed=0xC0FFEE)
    sw = [g.sweep(k)[0].copy() for k in range(10)]
    gpu, pipe = LoamGpu(), orc.Pipeline()
    res = []
    for k, xyz in enumerate(sw):
        r, o = gpu.process_sweep(xyz), pipe.process(xyz)
        res.append(r)
        assert (r.counts.n_full, r.counts.n_sharp, r.counts.n_less_sharp, r.counts.n_flat, r.counts.n_less_flat) == \
            (o.n_full, o.n_sharp, o.n_less_sharp, o.n_flat, o.n_less_flat), k
        assert np.array_equal(np.array(r.odom.transform_sum, np.float32), np.array(o.odom, np.float32)), k
        assert r.mapping_ran == o.mapping_ran
        if r.mapping_ran:
            assert np.array_equal(np.array(r.map.transform_aft_mapped, np.float32), np.array(o.mapped, np.float32)), k
            assert r.map.iterations == o.map_iters
    gpu.close()
    p = LoamGpuPipeline()
    for xyz in sw:
        p.submit(xyz)
    for k in range(len(sw)):
        r = p.wait()
        assert list(r.odom.transform_sum) == list(res[k].odom.transform_sum), k
        assert list(r.map.transform_aft_mapped) == list(res[k].map.transform_aft_mapped), k
    p.close()


def test_pipeline_with_pose_message_hop(orc, sweeps16):
    """pose_message_hop = 1: the odometry pose reaches mapping through the quaternion message (LO:1066-1078 -> LM:322-332) like
    between the reference's nodes; equal, bit for bit, to the oracle wired the same way -- fused call and pipelined mode."""
    from gpscalibration_b200 import LoamGpu, LoamGpuPipeline
    gpu, pipe = LoamGpu(pose_message_hop=True), orc.Pipeline()
    pipe.set_ros_hop(True)
    res = []
    for k, xyz in enumerate(sweeps16[:14]):
        r, o = gpu.process_sweep(xyz), pipe.process(xyz)
        res.append(r)
        assert np.array_equal(np.array(r.odom.transform_sum, np.float32), np.array(o.odom, np.float32)), k
        assert r.mapping_ran == o.mapping_ran
        if r.mapping_ran:
            assert np.array_equal(np.array(r.map.transform_aft_mapped, np.float32), np.array(o.mapped, np.float32)), k
            assert r.map.iterations == o.map_iters
    gpu.close()
    p = LoamGpuPipeline(pose_message_hop=True)
    for xyz in sweeps16[:14]:
        p.submit(xyz)
    for k in range(14):
        r = p.wait()
        assert list(r.map.transform_aft_mapped) == list(res[k].map.transform_aft_mapped), k
    p.close()


# ------------------------------------------------------------------------------------------------ odometry (a6-a12)
def _odom_pair(orc, sweeps16, a=0, b=1):
    sr = orc.ScanRegistration()
    f0 = sr.extract(sweeps16[a])
    f1 = sr.extract(sweeps16[b])
    T0 = np.zeros(6, np.float32)
    return f1, orc.transform_to_end(f0["less_sharp"], T0), orc.transform_to_end(f0["less_flat"], T0)


def test_transform_to_end_parity(gpu, orc, sweeps16):
    f1, _, _ = _odom_pair(orc, sweeps16)
    T = np.array([0.003, 0.021, -0.002, 0.05, -0.01, 1.02], np.float32)
    imu = np.array([0.01, -0.02, 0.005, 0.012, -0.018, 0.004, 0.1, -0.05, 0.02, 0.3, 0.1, -0.2], np.float32)
    for imu_t in (None, imu):
        ref = orc.transform_to_end(f1["less_flat"], T, imu_t)
        got = gpu.transform_to_end(f1["less_flat"], T, imu_t)
        assert np.array_equal(got[:, 3], ref[:, 3])
        # sin/cos of the per-point angles: device port of glibc sinf/cosf, bit-identical on a matching host libm
        assert np.array_equal(got[:, :3], ref[:, :3])


def test_odom_iterations_parity(gpu, orc, sweeps16):
    """cfg 1: one full Gauss-Newton registration of sweep 1 against sweep 0, iteration by iteration."""
    f1, last_c, last_s = _odom_pair(orc, sweeps16)
    gpu.odom_set_inputs(f1["sharp"], f1["flat"], last_c, last_s)
    oi = orc.OdomIter(f1["sharp"], f1["flat"], last_c, last_s, brute=True)
    ns, nf = f1["sharp"].shape[0], f1["flat"].shape[0]
    T = np.zeros(6, np.float32)
    state = np.zeros(37, np.float32)
    worst = 0.0
    for it in range(25):
        rAtA, rAtB, rn = oi.iterate(it, T)
        AtA, AtB, n = gpu.odom_iter(it, T)
        assert n == rn, (it, n, rn)
        if it % 5 == 0:
            c1, c2, s1, s2, s3 = gpu.odom_corr(ns, nf)
            assert np.array_equal(c1, oi.c1) and np.array_equal(c2, oi.c2), it  # bit-exact correspondences
            assert np.array_equal(s1, oi.s1) and np.array_equal(s2, oi.s2) and np.array_equal(s3, oi.s3), it
        if rn < 10:
            continue
        sa, sb = np.abs(rAtA).max(), np.abs(rAtB).max()
        worst = max(worst, np.abs(AtA - rAtA).max() / sa, np.abs(AtB - rAtB).max() / sb)
        assert np.abs(AtA - rAtA).max() <= 1e-5 * sa and np.abs(AtB - rAtB).max() <= 1e-5 * sb, it
        X = orc.gn_solve(rAtA, rAtB, it, 10.0, state)
        T = (T + X).astype(np.float32)
    assert abs(T[5]) > 0.5  # the registration actually moved (1 m/sweep forward)
    print("odom worst rel err", worst)


def test_host_gn_solve_matches_oracle(orc):
    from gpscalibration_b200 import capi
    rng = np.random.default_rng(5)
    for trial in range(20):
        A = rng.normal(size=(200, 6)).astype(np.float32)
        if trial % 3 == 0:
            A[:, 4] *= 1e-3  # near-degenerate direction -> projection path
        AtA = (A.T.astype(np.float64) @ A.astype(np.float64)).astype(np.float32)
        AtB = rng.normal(size=6).astype(np.float32)
        s1, s2 = np.zeros(37, np.float32), np.zeros(37, np.float32)
        for it in (0, 1):
            X1 = capi.gn_solve(AtA, AtB, it, 10.0, s1)
            X2 = orc.gn_solve(AtA, AtB, it, 10.0, s2)
            assert np.array_equal(X1, X2) and np.array_equal(s1, s2)


# ------------------------------------------------------------------------------------------------ mapping (a13-a18)
def _run_pipeline(gpu, orc_pipe, sweeps, check=None):
    out = []
    for k, xyz in enumerate(sweeps):
        r = gpu.process_sweep(xyz)
        o = orc_pipe.process(xyz)
        out.append((r, o))
        if check:
            check(k, r, o)
    return out


def test_map_iteration_parity(gpu, orc, sweeps16):
    """Correspondences (5-NN) bit-exact and normal equations within 1e-5 on a map built by the pipeline itself."""
    pipe = orc.Pipeline()
    _run_pipeline(gpu, pipe, sweeps16[:10])
    cs, ss = gpu.cloud("corner_stack"), gpu.cloud("surf_stack")
    cm, sm = gpu.cloud("corner_map"), gpu.cloud("surf_map")
    assert cm.shape[0] > 1000 and sm.shape[0] > 5000
    T = np.array([0.001, 0.17, -0.002, 0.6, 0.05, 8.5], np.float32)
    rAtA, rAtB, rn, rcc, rcs = orc.map_iteration(cs, ss, cm, sm, T, brute=False)
    gpu.map_set_inputs(cs, ss, cm, sm)
    AtA, AtB, n = gpu.map_iter(0, T)
    cc, cs5 = gpu.map_corr(cs.shape[0], ss.shape[0])
    assert np.array_equal(cc, rcc) and np.array_equal(cs5, rcs)
    assert (rcs[:, 0] >= 0).sum() > 1000
    assert n == rn
    sa, sb = np.abs(rAtA).max(), np.abs(rAtB).max()
    assert np.abs(AtA - rAtA).max() <= 1e-5 * sa and np.abs(AtB - rAtB).max() <= 1e-5 * sb


def test_pipeline_parity_vlp16(gpu, orc, sweeps16):
    """cfg 2 shape (short): extract -> odometry -> mapping over consecutive sweeps, poses vs the oracle."""
    pipe = orc.Pipeline()
    worst = {"t": 0.0, "r": 0.0}

    def check(k, r, o):
        assert (r.counts.n_full, r.counts.n_sharp, r.counts.n_less_sharp, r.counts.n_flat, r.counts.n_less_flat) == \
            (o.n_full, o.n_sharp, o.n_less_sharp, o.n_flat, o.n_less_flat), k
        assert r.odom.odom_published == o.odom_published and r.mapping_ran == o.mapping_ran, k
        go, ro = np.array(r.odom.transform_sum), np.array(o.odom)
        assert np.abs(go[:3] - ro[:3]).max() <= 1e-5 and np.abs(go[3:] - ro[3:]).max() <= 1e-4, (k, go, ro)
        assert r.odom.iterations == o.odom_iters, (k, r.odom.iterations, o.odom_iters)
        worst["r"] = max(worst["r"], np.abs(go[:3] - ro[:3]).max())
        worst["t"] = max(worst["t"], np.abs(go[3:] - ro[3:]).max())
        if r.mapping_ran:
            assert (r.map.n_corner_stack, r.map.n_surf_stack, r.map.n_corner_map, r.map.n_surf_map) == \
                (o.n_corner_stack, o.n_surf_stack, o.n_corner_map, o.n_surf_map), k
            gm, rm = np.array(r.map.transform_aft_mapped), np.array(o.mapped)
            assert np.abs(gm[:3] - rm[:3]).max() <= 1e-5 and np.abs(gm[3:] - rm[3:]).max() <= 1e-4, (k, gm, rm)
            assert r.map.iterations == o.map_iters, (k, r.map.iterations, o.map_iters)
            worst["map"] = max(worst.get("map", 0.0), float(np.abs(gm - rm).max()))

    _run_pipeline(gpu, pipe, sweeps16, check)
    print("pipeline worst pose diff", worst)


def test_pipeline_parity_hdl64(orc):
    """cfg 3 shape: 64 rings, ~120 k points, ~18 k features per sweep.  Exercises what the VLP-16 sequence does not: several
    rows per thread in the device Gauss-Newton loop, the multi-pass radix sort (> 16 k pairs), the split / big voxel paths.
    Poses, iteration counts and cloud sizes equal to the oracle, bit for bit."""
    from gpscalibration_b200 import LoamGpu, SweepGenerator
    g = SweepGenerator(sensor=2, scene=1, seed=0xC0FFEE)
    step = 26.8 / 63.0
    gpu = LoamGpu(n_scans=64, ring_mode=1, ring_ang_min=-24.8, ring_ang_step=step)
    pipe = orc.Pipeline(64, 1, -24.8, step)
    seen_map = 0
    for k in range(7):
        xyz = g.sweep(k)[0]
        r, o = gpu.process_sweep(xyz), pipe.process(xyz)
        assert (r.counts.n_full, r.counts.n_sharp, r.counts.n_less_sharp, r.counts.n_flat, r.counts.n_less_flat) == \
            (o.n_full, o.n_sharp, o.n_less_sharp, o.n_flat, o.n_less_flat), k
        assert r.odom.odom_published == o.odom_published and r.mapping_ran == o.mapping_ran, k
        assert np.array_equal(np.array(r.odom.transform_sum, np.float32), np.array(o.odom, np.float32)), k
        if o.odom_published:
            assert r.odom.iterations == o.odom_iters, (k, r.odom.iterations, o.odom_iters)
        if r.mapping_ran:
            seen_map += 1
            assert (r.map.n_corner_stack, r.map.n_surf_stack, r.map.n_corner_map, r.map.n_surf_map) == \
                (o.n_corner_stack, o.n_surf_stack, o.n_corner_map, o.n_surf_map), k
            assert np.array_equal(np.array(r.map.transform_aft_mapped, np.float32), np.array(o.mapped, np.float32)), k
            assert r.map.iterations == o.map_iters, k
    assert r.counts.n_sharp + r.counts.n_flat > 8192 and seen_map >= 2
    gpu.close()


def test_pipeline_reset_protocol(gpu, orc, sweeps16):
    """IMControl{false}: odometry re-initialises, mapping resets on the zero pose."""
    pipe = orc.Pipeline()
    for k in range(5):
        gpu.process_sweep(sweeps16[k])
        pipe.process(sweeps16[k])
    gpu.reset()
    pipe.reset()
    for k in range(5, 10):
        r = gpu.process_sweep(sweeps16[k])
        o = pipe.process(sweeps16[k])
        assert r.odom.odom_published == o.odom_published and r.mapping_ran == o.mapping_ran
        assert np.abs(np.array(r.odom.transform_sum) - np.array(o.odom)).max() <= 1e-4
        if r.mapping_ran:
            assert (r.map.n_corner_map, r.map.n_surf_map) == (o.n_corner_map, o.n_surf_map)


def test_nodewise_equals_fused(orc, sweeps16):
    from gpscalibration_b200 import LoamPipeline
    a, b = LoamPipeline(), LoamPipeline()
    for k in range(6):
        r = a.process(sweeps16[k])
        o, m = b.process_nodewise(sweeps16[k])
        assert np.array_equal(np.array(r.odom.transform_sum), np.array(o.transform_sum))
        assert (m is not None) == bool(r.mapping_ran)
        if m is not None:
            assert np.array_equal(np.array(r.map.transform_aft_mapped), np.array(m.transform_aft_mapped))


def test_registered_and_surround_clouds(orc, sweeps16):
    from gpscalibration_b200 import LoamGpu
    gpu = LoamGpu(want_registered=True, want_surround=True)
    pipe = orc.Pipeline(keep_clouds=True)
    seen_surround = 0
    for k in range(14):
        r = gpu.process_sweep(sweeps16[k])
        o = pipe.process(sweeps16[k])
        if r.mapping_ran:
            reg, rreg = gpu.cloud("registered"), pipe.cloud("registered")
            assert reg.shape == rreg.shape and np.array_equal(reg.view(np.uint32), rreg.view(np.uint32)), k
            if r.map.surround_published:
                s, rs = gpu.cloud("surround"), pipe.cloud("surround")
                assert s.shape == rs.shape and np.array_equal(s.view(np.uint32), rs.view(np.uint32)), k
                seen_surround += 1
    assert seen_surround >= 2
    gpu.close()


def test_pipelined_mode_equals_fused(sweeps16):
    """loam_pipeline_* (three stage threads, SR | LO | LM) must give exactly the results of loam_process_sweep."""
    from gpscalibration_b200 import LoamGpu, LoamGpuPipeline
    a = LoamGpu()
    ref = [a.process_sweep(x) for x in sweeps16]
    a.close()
    p = LoamGpuPipeline()
    for rep in range(2):  # second pass after an ordered reset
        for x in sweeps16:
            p.submit(x)
        got = [p.wait() for _ in sweeps16]
        assert p.pending == 0
        if rep == 0:
            for k, (r, g) in enumerate(zip(ref, got)):
                assert (r.counts.n_full, r.counts.n_less_flat) == (g.counts.n_full, g.counts.n_less_flat), k
                assert list(r.odom.transform_sum) == list(g.odom.transform_sum), k
                assert r.odom.iterations == g.odom.iterations and r.mapping_ran == g.mapping_ran, k
                if r.mapping_ran:
                    assert list(r.map.transform_aft_mapped) == list(g.map.transform_aft_mapped), k
                    assert (r.map.n_corner_map, r.map.n_surf_map, r.map.iterations) == (g.map.n_corner_map, g.map.n_surf_map, g.map.iterations)
        p.reset()
    p.close()


def test_pipelined_surround_cloud_equals_blocking_call():
    """want_surround in pipelined mode: the cloud is finished on the pipeline's output stage (gather enqueued by the mapping
    stage, voxel grid + count on the output thread) -- count of every surround run and the last cloud must equal what the
    blocking call (which does LM:1081-1101 in line) produces, over enough sweeps for several surround runs and a reset."""
    from gpscalibration_b200 import LoamGpu, LoamGpuPipeline, SweepGenerator
    gen = SweepGenerator(sensor=0, scene=0, seed=0xC0FFEE)
    sweeps = [gen.sweep(k)[0].copy() for k in range(44)]
    a = LoamGpu(want_registered=True, want_surround=True)
    p = LoamGpuPipeline(want_registered=True, want_surround=True)
    for rep in range(2):
        a.reset()
        ref, last_cloud = [], None
        for x in sweeps:
            r = a.process_sweep(x)
            ref.append((r.mapping_ran, r.map.surround_published, r.map.n_surround, r.map.n_registered, list(r.map.transform_aft_mapped)))
            if r.mapping_ran and r.map.surround_published:
                last_cloud = a.cloud("surround")
        p.reset()
        for x in sweeps:
            p.submit(x)
        got = [p.wait() for _ in sweeps]
        assert p.pending == 0
        n_sur = 0
        for k, (r, g) in enumerate(zip(ref, got)):
            assert r == (g.mapping_ran, g.map.surround_published, g.map.n_surround, g.map.n_registered, list(g.map.transform_aft_mapped)), (rep, k)
            n_sur += int(bool(g.mapping_ran and g.map.surround_published))
        assert n_sur >= 4 and all(r[2] > 0 for r in ref if r[0] and r[1])
        cloud = p.output_cloud("surround")
        assert cloud.shape == last_cloud.shape and np.array_equal(cloud.view(np.uint32), last_cloud.view(np.uint32))
    a.close()
    p.close()


def test_mapping_cube_grid_rolls_like_the_reference(orc, sweeps16):
    """K11 / a14 / a18: drive laserMapping alone with odometry poses that travel hundreds of metres (and back), so the
    21 x 11 x 21 cube grid re-centres in every direction (LM:497-657), cubes get cleared, points land in cubes outside
    the voxel-gridded set and the arena compacts.  Geometry is meaningless here (the same two clouds every time); what
    must hold is bit-equality with the oracle's cube bookkeeping, gather order and voxel grids."""
    from gpscalibration_b200 import LoamGpu
    f = orc.ScanRegistration().extract(sweeps16[0])
    T0 = np.zeros(6, np.float32)
    corner = orc.transform_to_end(f["less_sharp"], T0)
    surf = orc.transform_to_end(f["less_flat"], T0)
    gpu = LoamGpu()
    lm = orc.LaserMapping()
    gpu.odom_set_inputs(f["sharp"], f["flat"], corner, surf)  # fills corner_last / surf_last of the handle
    path = [(0.0, 0.0, 0.0)]
    for step in ((45.0, 0.0, 0.0),) * 10 + ((0.0, 0.0, 45.0),) * 9 + ((0.0, 40.0, 0.0),) * 5 + ((-45.0, -20.0, -45.0),) * 14:
        path.append(tuple(a + b for a, b in zip(path[-1], step)))
    for k, (x, y, z) in enumerate(path[1:]):
        Tsum = np.array([0.001 * k, 0.02 * k, -0.0005 * k, x, y, z], np.float32)
        ro = lm.step(Tsum, corner, surf)
        gpu.mapping_odometry(Tsum)
        r = gpu.mapping_process()
        assert (r.n_corner_map, r.n_surf_map, r.n_corner_stack, r.n_surf_stack, r.iterations) == \
            (int(ro[19]), int(ro[20]), int(ro[21]), int(ro[22]), int(ro[18])), (k, x, y, z)
        assert np.array_equal(np.array(r.transform_aft_mapped, np.float32), ro[:6]), k
        assert np.array_equal(np.array(r.transform_tobe_mapped, np.float32), ro[12:18]), k
    gpu.close()


# ------------------------------------------------------------------------------------------------ PointCloud2 wire (N3)
@pytest.mark.gpu
def test_pointcloud2_wire_in_and_out(gpu, sweeps16):
    """the library reads sensor_msgs/PointCloud2 payloads directly (any point_step, any alignment) and
    emits the payload pcl::toROSMsg would (SR:260-261, 689-726)."""
    xyz = sweeps16[2]
    n = xyz.shape[0]
    ref_counts = gpu.extract(xyz)
    ref = {nm: gpu.cloud(nm) for nm in ("full", "sharp", "less_sharp", "flat", "less_flat")}
    # Velodyne driver layout (PointXYZIR, point_step 22) behind a 3-byte header offset: nothing is 4-byte aligned
    rec = np.zeros((n, 22), np.uint8)
    rec[:, 0:12] = xyz.view(np.uint8).reshape(n, 12)
    rec[:, 16:20] = np.full(n, 7.0, np.float32).view(np.uint8).reshape(n, 4)
    rec[:, 20:22] = (np.arange(n) % 16).astype(np.uint16).view(np.uint8).reshape(n, 2)
    blob = np.concatenate([np.zeros(3, np.uint8), rec.reshape(-1)])
    c = gpu.extract_wire(blob[3:], 22)
    assert [c.n_full, c.n_sharp, c.n_less_sharp, c.n_flat, c.n_less_flat] == \
        [ref_counts.n_full, ref_counts.n_sharp, ref_counts.n_less_sharp, ref_counts.n_flat, ref_counts.n_less_flat]
    for nm, want in ref.items():
        assert np.array_equal(gpu.cloud(nm), want), nm
    # PCL PointXYZI layout (point_step 32)
    rec32 = np.zeros((n, 8), np.float32)
    rec32[:, 0:3] = xyz
    rec32[:, 3] = 1.0
    c = gpu.extract_wire(rec32.view(np.uint8), 32)
    assert c.n_full == ref_counts.n_full and np.array_equal(gpu.cloud("less_flat"), ref["less_flat"])
    # and out again
    for nm, want in ref.items():
        wire = gpu.cloud_wire(nm)
        assert wire.shape == (want.shape[0], 32)
        f = wire.view(np.float32).reshape(-1, 8)
        assert np.array_equal(f[:, 0:3], want[:, 0:3]) and np.array_equal(f[:, 4], want[:, 3])
        assert np.all(f[:, 3] == 1.0) and not f[:, 5:].any()
