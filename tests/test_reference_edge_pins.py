"""Pins the oracle on the EDGE inputs and flows the GPU parity tests use -- for scanRegistration -- empty rings (true VLP-16 angles, sky-facing
top rings, an occluded ring mid-sequence: the stale / overlapping scanStartInd / scanEndInd of SR:480-490 and the five
never re-initialised entries of the static arrays, state carried from sweep to sweep), NaN / inf points and a ragged tail,
dense rings with curvature ties -- against the reference's OWN scanRegistration.cpp (a private copy of
oracle/_ref/libref_sr.so per sequence: the node keeps its state in file-scope globals), bit for bit on all five clouds.
Further down: the whole pipeline on empty-ring sweeps, the cube grid rolling over hundreds of metres, and the reset protocol,
against private copies of all three nodes.  The GPU tests compare the CUDA path with the oracle on the same inputs; this
closes the chain to the reference's code."""
import numpy as np
import pytest

NAMES = ("full", "sharp", "less_sharp", "flat", "less_flat")


def _pin(orc, sweeps):
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    r = ref.SrWithImu()  # no IMU message is ever sent: imuPointerLast stays -1 and the plain path runs (SR:364)
    o = orc.ScanRegistration()
    sizes = []
    for k, xyz in enumerate(sweeps):
        feat, tr = r.process(xyz, 100.0 + 0.1 * k)
        oc = o.extract(xyz)
        assert not tr.any()
        for i, nm in enumerate(NAMES):
            a, b = feat[i], oc[nm]
            assert a.shape == b.shape, (k, nm, a.shape, b.shape)
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (k, nm)
        sizes.append([f.shape[0] for f in feat])
    return o, sizes


def _elev(xyz):
    return np.degrees(np.arctan2(xyz[:, 2], np.hypot(xyz[:, 0], xyz[:, 1])))


def test_true_vlp16_angles_four_sweeps(orc):
    from gpscalibration_b200 import SweepGenerator
    g = SweepGenerator(sensor=1, scene=0, seed=0xC0FFEE)
    o, sizes = _pin(orc, [g.sweep(k)[0].copy() for k in range(4)])
    assert (o.ints("scan_start")[[6, 8, 10]] == 0).all() and (o.ints("scan_end")[[5, 7, 9]] == 0).all()
    assert sizes[-1][4] > sizes[-1][0] // 4  # the virtual rings contribute the earlier rings again


def test_nan_inf_and_ragged_tail(orc, sweeps16):
    xyz = sweeps16[2].copy()
    rng = np.random.default_rng(3)
    bad = rng.choice(xyz.shape[0], 500, replace=False)
    xyz[bad[:250], 0] = np.nan
    xyz[bad[250:], 2] = np.inf
    xyz[0] = np.nan
    xyz[-1] = np.nan
    _pin(orc, [xyz[:-777]])


def test_top_rings_missing_then_full(orc, sweeps16):
    cuts = [sweeps16[k][_elev(sweeps16[k]) < 4.0] for k in (3, 4)]
    o, _ = _pin(orc, cuts[:1])
    assert o.ints("scan_start")[15] == 0 and o.ints("scan_end")[12] == 0
    _pin(orc, cuts + [sweeps16[5]])  # a fully populated sweep afterwards: the five stale entries keep their marks


def test_occluded_rings_mid_sequence(orc, sweeps16):
    xyz = sweeps16[1]
    e = _elev(xyz)
    _pin(orc, [sweeps16[0], xyz[np.abs(e + 2.0) > 0.5], xyz[e > -14.0], sweeps16[2]])


def test_dense_rings_with_curvature_ties(orc):
    from test_select_rounds import _dense_cylinder
    _pin(orc, [_dense_cylinder()])


def test_whole_pipeline_on_true_vlp16_angles(orc):
    """Empty rings every sweep, through all three nodes: feature clouds, odometry and mapped poses, publish pattern and map
    sizes of the oracle's pipeline equal to the reference's own three translation units (private copies), bit for bit."""
    from gpscalibration_b200 import SweepGenerator
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    g = SweepGenerator(sensor=1, scene=0, seed=0xC0FFEE)
    nodes = ref.PrivateNodes()
    pipe = orc.Pipeline()
    pipe.set_ros_hop(True)
    try:
        ran = 0
        for k in range(10):
            x = g.sweep(k)[0].copy()
            r = nodes.process(x, 300.0 + 0.1 * k)
            o = pipe.process(x)
            for i, nm in enumerate(NAMES):
                assert np.array_equal(r.features[i].view(np.uint32), pipe.cloud(nm).view(np.uint32)), (k, nm)
            assert r.odom_published == bool(o.odom_published) and r.mapping_ran == bool(o.mapping_ran), k
            assert np.array_equal(r.odom, np.array(o.odom, np.float32)), k
            if r.mapping_ran:
                ran += 1
                assert np.array_equal(r.mapped, np.array(o.mapped, np.float32)), k
        assert ran >= 4 and list(nodes.map_size()) == list(pipe.map_size())
    finally:
        nodes.close()


def test_mapping_cube_grid_rolls_like_the_reference_code(orc, sweeps16):
    """laserMapping alone, driven along a path that travels hundreds of metres and back: the 21 x 11 x 21 cube grid re-centres in
    every direction (LM:497-657), cubes are cleared, points land in cubes outside the voxel-gridded set.  The oracle's node
    against the reference's own laserMapping.cpp (private copy) on the SAME parsed transformSum (read back from the
    reference's handler, so the quaternion message hop is not part of this test): transformAftMapped / BefMapped /
    TobeMapped and the map sizes, bit for bit.  The GPU test of the same name compares the CUDA path with the oracle."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    f = orc.ScanRegistration().extract(sweeps16[0])
    T0 = np.zeros(6, np.float32)
    corner = orc.transform_to_end(f["less_sharp"], T0)
    surf = orc.transform_to_end(f["less_flat"], T0)
    full = orc.transform_to_end(f["full"], T0)
    nodes = ref.PrivateNodes()
    lm = orc.LaserMapping()
    path = [(0.0, 0.0, 0.0)]
    for step in ((45.0, 0.0, 0.0),) * 10 + ((0.0, 0.0, 45.0),) * 9 + ((0.0, 40.0, 0.0),) * 5 + ((-45.0, -20.0, -45.0),) * 14:
        path.append(tuple(a + b for a, b in zip(path[-1], step)))
    try:
        for k, (x, y, z) in enumerate(path[1:]):
            Tsum = np.array([0.001 * k, 0.02 * k, -0.0005 * k, x, y, z], np.float32)
            m = np.zeros(24, np.float32)
            pose7 = ref.pose7_from_T(Tsum)
            nodes.lm.ref_lm_step(corner.ctypes.data, corner.shape[0], surf.ctypes.data, surf.shape[0], full.ctypes.data, full.shape[0],
                                 pose7.ctypes.data, 400.0 + 0.1 * k, m.ctypes.data)
            ro = lm.step(m[18:24].copy(), corner, surf, full)  # the transformSum the reference's handler parsed
            assert np.array_equal(m[:18].view(np.uint32), ro[:18].view(np.uint32)), (k, x, y, z)
            assert list(nodes.map_size()) == [int(ro[23]), int(ro[24])], k
    finally:
        nodes.close()


def test_reset_protocol_like_the_reference_code(orc, sweeps16):
    """IMControl{systemInited=false} between two sweeps (IN:281-284): laserOdometry re-initialises on the next sweep, laserMapping
    resets when it sees the zero pose (LM:316-319).  The oracle's pipeline against private copies of the reference's three
    nodes across the reset: clouds, poses, publish pattern and map sizes, bit for bit."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    nodes = ref.PrivateNodes()
    pipe = orc.Pipeline()
    pipe.set_ros_hop(True)
    try:
        for k in range(11):
            if k == 5:
                nodes.control_reset()
                pipe.reset()
            r = nodes.process(sweeps16[k], 500.0 + 0.1 * k)
            o = pipe.process(sweeps16[k])
            for i, nm in enumerate(NAMES):
                assert np.array_equal(r.features[i].view(np.uint32), pipe.cloud(nm).view(np.uint32)), (k, nm)
            assert r.odom_published == bool(o.odom_published) and r.mapping_ran == bool(o.mapping_ran), k
            assert np.array_equal(r.odom, np.array(o.odom, np.float32)), k
            if r.mapping_ran:
                assert np.array_equal(r.mapped, np.array(o.mapped, np.float32)), k
            assert list(nodes.map_size()) == list(pipe.map_size()), k
        assert np.abs(r.odom[3:]).max() > 0.1  # and the odometry moved again after the reset
    finally:
        nodes.close()


def test_registered_and_surround_clouds_like_the_reference_code(orc, sweeps16):
    """/velodyne_cloud_registered (LM:1103-1112, every mapping run) and /laser_cloud_surround (LM:1081-1101, the first run and
    every fifth one after it): the oracle's clouds against what the reference's own laserMapping.cpp publishes, bit for bit.
    The GPU tests compare the CUDA path's two clouds with the oracle's."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    nodes = ref.PrivateNodes()
    pipe = orc.Pipeline()
    pipe.set_ros_hop(True)
    try:
        n_surround = 0
        for k in range(14):
            r = nodes.process(sweeps16[k], 600.0 + 0.1 * k)
            o = pipe.process(sweeps16[k])
            assert r.mapping_ran == bool(o.mapping_ran), k
            if not r.mapping_ran:
                continue
            a, b = nodes.lm_cloud(1), pipe.cloud("registered")
            assert a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32)), k
            s = pipe.cloud("surround")  # empty unless this run published it
            if s.shape[0]:
                n_surround += 1
                c = nodes.lm_cloud(0)
                assert c.shape == s.shape and np.array_equal(c.view(np.uint32), s.view(np.uint32)), k
        assert n_surround == 2
    finally:
        nodes.close()


@pytest.mark.parametrize("scene,seed", [(1, 0xC0FFEE + 7), (0, 12345)])
def test_whole_pipeline_other_scenes_and_seeds(orc, scene, seed):
    """The three nodes on sequences other than the golden one (the second synthetic scene; another seed): the oracle's pipeline
    against private copies of the reference's own translation units, bit for bit over twelve sweeps."""
    from gpscalibration_b200 import SweepGenerator
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    g = SweepGenerator(sensor=0, scene=scene, seed=seed)
    nodes = ref.PrivateNodes()
    pipe = orc.Pipeline()
    pipe.set_ros_hop(True)
    try:
        for k in range(12):
            x = g.sweep(k)[0].copy()
            r = nodes.process(x, 700.0 + 0.1 * k)
            o = pipe.process(x)
            for i, nm in enumerate(NAMES):
                assert np.array_equal(r.features[i].view(np.uint32), pipe.cloud(nm).view(np.uint32)), (k, nm)
            assert r.odom_published == bool(o.odom_published) and r.mapping_ran == bool(o.mapping_ran), k
            assert np.array_equal(r.odom, np.array(o.odom, np.float32)), k
            assert np.array_equal(r.rel, np.array(o.rel, np.float32)), k
            if r.mapping_ran:
                assert np.array_equal(r.mapped, np.array(o.mapped, np.float32)), k
        assert list(nodes.map_size()) == list(pipe.map_size())
    finally:
        nodes.close()
