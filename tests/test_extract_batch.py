"""SURVEY §8b `*_batch`: loam_extract_batch launches every extraction kernel once for B sequences (grid.y = sequence).
Its results must be those of B loam_extract calls bit for bit -- counts and all five clouds -- for ragged batches (different
point counts, NaNs, an empty sweep, a sweep with empty rings that takes the serial replay), over several sweeps per
sequence (the per-sequence state between sweeps stays separate), and the features must feed the unbatched odometry /
mapping of each sequence unchanged."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CLOUDS = ("full", "sharp", "less_sharp", "flat", "less_flat")


def _counts(c):
    return (c.n_full, c.n_sharp, c.n_less_sharp, c.n_flat, c.n_less_flat)


def _sequences(B, n_sweeps):
    from gpscalibration_b200 import SweepGenerator
    seqs = []
    for b in range(B):
        sensor = 1 if b == 2 else 0  # member 2: true VLP-16 angles -> rings 6 / 8 / 10 of the reference's table stay empty
        g = SweepGenerator(sensor=sensor, scene=b % 2, seed=0xC0FFEE + 1000 * b, t_offset=37.0 * b)
        sw = [g.sweep(k)[0].copy() for k in range(n_sweeps)]
        if b == 1:  # ragged + invalid points
            sw = [x[: x.shape[0] - 500 * (k + 1)].copy() for k, x in enumerate(sw)]
            sw[0][::97, 1] = np.nan
        seqs.append(sw)
    return seqs


@pytest.mark.parametrize("B", [1, 3, 8])
def test_batched_extraction_equals_single_calls(B):
    from gpscalibration_b200 import LoamGpu, capi
    n_sweeps = 3
    seqs = _sequences(B, n_sweeps)
    single = [LoamGpu() for _ in range(B)]
    batch = [LoamGpu() for _ in range(B)]
    for k in range(n_sweeps):
        sweeps = [seqs[b][k] for b in range(B)]
        if B >= 3 and k == 1:
            sweeps[0] = sweeps[0][:0]  # an empty sweep inside the batch
        ref = [single[b].extract(sweeps[b]) for b in range(B)]
        got = capi.extract_batch(batch, sweeps)
        for b in range(B):
            assert _counts(got[b]) == _counts(ref[b]), (k, b)
            for name in CLOUDS:
                a, c = single[b].cloud(name), batch[b].cloud(name)
                assert a.shape == c.shape and np.array_equal(a.view(np.uint32), c.view(np.uint32)), (k, b, name)
    launches_single = sum(h.stats()["launches"] for h in single)
    launches_batch = sum(h.stats()["launches"] for h in batch)
    if B == 8:
        assert launches_batch < launches_single / 3  # the point of batching (the fall-back members launch on their own)
    for h in single + batch:
        h.close()


def test_batched_features_drive_the_unbatched_pipeline():
    """Extraction batched over four sequences, odometry + mapping per sequence: poses equal to loam_process_sweep."""
    from gpscalibration_b200 import LoamGpu, capi
    B, n_sweeps = 4, 12
    seqs = _sequences(B, n_sweeps)
    ref = [LoamGpu() for _ in range(B)]
    bat = [LoamGpu() for _ in range(B)]
    for k in range(n_sweeps):
        capi.extract_batch(bat, [seqs[b][k] for b in range(B)])
        for b in range(B):
            r = ref[b].process_sweep(seqs[b][k])
            o = bat[b].odometry_process()
            assert list(o.transform_sum) == list(r.odom.transform_sum) and o.iterations == r.odom.iterations, (k, b)
            if o.odom_published:
                bat[b].mapping_odometry(np.array(o.transform_sum, np.float32))
            if o.odom_published and o.fullres_published:
                m = bat[b].mapping_process()
                assert r.mapping_ran and list(m.transform_aft_mapped) == list(r.map.transform_aft_mapped), (k, b)
    for h in ref + bat:
        h.close()


def test_batch_rejects_bad_arguments():
    from gpscalibration_b200 import LoamGpu, capi
    a, b = LoamGpu(), LoamGpu(n_scans=64, ring_mode=1, ring_ang_min=-24.8, ring_ang_step=26.8 / 63.0)
    x = np.zeros((10, 3), np.float32)
    with pytest.raises(capi.LoamError):
        capi.extract_batch([a, b], [x, x])  # different ring tables in one batch
    with pytest.raises(capi.LoamError):
        capi.extract_batch([a, a], [x, x])  # the same handle twice
    a.close()
    b.close()


@pytest.mark.parametrize("lockstep", [False, True])
def test_pipelines_fed_by_batched_extraction_equal_plain_pipelines(lockstep):
    """loam_pipeline_submit_batch (extraction batched) / loam_pipeline_submit_lockstep (extraction + odometry batched in the
    caller's thread): four pipelines, mapping / output stages per pipeline -- every result equal to the same pipelines fed
    with loam_pipeline_submit, across a reset."""
    from gpscalibration_b200 import LoamGpuPipeline, capi
    B, n_sweeps = 4, 30
    seqs = _sequences(B, n_sweeps)

    def key(r):
        t = (r.counts.n_full, r.counts.n_less_flat, list(r.odom.transform_sum), r.odom.iterations, r.mapping_ran)
        if r.mapping_ran:
            t += (list(r.map.transform_aft_mapped), r.map.iterations, r.map.n_corner_map, r.map.n_surf_map, r.map.n_surround, r.map.n_registered)
        return t

    plain = [LoamGpuPipeline(want_registered=True, want_surround=True) for _ in range(B)]
    # gn_max_ctas: a smaller cooperative grid for the mapping loop (several sequences side by side) must not change a result
    batched = [LoamGpuPipeline(want_registered=True, want_surround=True, gn_max_ctas=20) for _ in range(B)]
    ref = [[] for _ in range(B)]
    got = [[] for _ in range(B)]
    for k in range(n_sweeps):
        if k == 17:
            for p in plain + batched:
                p.reset()
        for b in range(B):
            plain[b].submit(seqs[b][k])
        capi.pipeline_submit_batch(batched, [seqs[b][k] for b in range(B)], lockstep=lockstep)
        if k >= 4:
            for b in range(B):
                ref[b].append(key(plain[b].wait()))
                got[b].append(key(batched[b].wait()))
    for b in range(B):
        while plain[b].pending:
            ref[b].append(key(plain[b].wait()))
        while batched[b].pending:
            got[b].append(key(batched[b].wait()))
        assert len(ref[b]) == n_sweeps and ref[b] == got[b], b
    for p in plain + batched:
        p.close()


@pytest.mark.parametrize("B", [1, 5])
def test_lockstep_odometry_batch_equals_single_calls(B):
    """loam_extract_batch + loam_odometry_process_batch (lock-step rounds, one launch per kernel for the batch) + per-handle
    mapping against loam_process_sweep per sequence: every pose, iteration count and flag equal, over enough sweeps for
    members to converge in different rounds, a reset of one member in mid-stream and a member with empty rings."""
    from gpscalibration_b200 import LoamGpu, capi
    n_sweeps = 16
    seqs = _sequences(B, n_sweeps)
    ref = [LoamGpu() for _ in range(B)]
    bat = [LoamGpu() for _ in range(B)]
    for k in range(n_sweeps):
        if k == 9 and B > 1:
            ref[1].reset()
            bat[1].reset()
        capi.extract_batch(bat, [seqs[b][k] for b in range(B)])
        outs = capi.odometry_process_batch(bat)
        for b in range(B):
            r = ref[b].process_sweep(seqs[b][k])
            o = outs[b]
            assert list(o.transform_sum) == list(r.odom.transform_sum) and list(o.transformation) == list(r.odom.transformation), (k, b)
            assert (o.iterations, o.odom_published, o.clouds_published, o.fullres_published, o.n_corner_last, o.n_surf_last) == \
                (r.odom.iterations, r.odom.odom_published, r.odom.clouds_published, r.odom.fullres_published, r.odom.n_corner_last,
                 r.odom.n_surf_last), (k, b)
            if o.odom_published:
                bat[b].mapping_odometry(np.array(o.transform_sum, np.float32))
            if o.odom_published and o.fullres_published:
                m = bat[b].mapping_process()
                assert r.mapping_ran and list(m.transform_aft_mapped) == list(r.map.transform_aft_mapped) and m.iterations == r.map.iterations, (k, b)
    if B > 1:
        assert sum(h.stats()["launches"] for h in bat) < 0.8 * sum(h.stats()["launches"] for h in ref)
    for h in ref + bat:
        h.close()
