"""SURVEY 8e (2), "optimised" variant: the 28-double all-reduce fused into the reduction kernel (peer stores + flags).

Two shards of one map, two handles ("ranks") driven by two host threads on ONE GPU: their reduction kernels exchange the
partial sums through each other's exchange buffers exactly as two processes would over NVLink (the only difference is
how the buffers are mapped: by pointer inside a process, through CUDA IPC between processes -- tools/bench_sharded.py
--fused-allreduce under torchrun covers that).  The global sums must equal the un-sharded iteration's.
"""
import threading

import numpy as np
import pytest


@pytest.mark.gpu
def test_fused_allreduce_two_ranks_one_gpu(orc, sweeps16):
    from gpscalibration_b200 import LoamGpu, sharding
    whole = LoamGpu()
    for xyz in sweeps16[:10]:
        whole.process_sweep(xyz)
    cs, ss = whole.cloud("corner_stack"), whole.cloud("surf_stack")
    cm, sm = whole.cloud("corner_map"), whole.cloud("surf_map")
    T = np.array([0.001, 0.17, -0.002, 0.6, 0.05, 8.5], np.float32)
    whole.map_set_inputs(cs, ss, cm, sm)
    want = [whole.map_iter(it, T) for it in range(3)]
    lo, hi = float(min(cm[:, 0].min(), sm[:, 0].min())), float(max(cm[:, 0].max(), sm[:, 0].max()))
    edges = sharding.slab_edges(lo - 1.0, hi + 1.0, 2)
    ranks = [LoamGpu(), LoamGpu()]
    handles = [g.shard_export() for g in ranks]
    got = [[None] * 3, [None] * 3]
    nq = 0
    for r, g in enumerate(ranks):
        g.shard_connect(handles, r)
        my_cs, my_ss = sharding.route_queries(cs, T, edges, r), sharding.route_queries(ss, T, edges, r)
        nq += my_cs.shape[0] + my_ss.shape[0]
        g.map_set_inputs(my_cs, my_ss, sharding.shard_map(cm, edges, r), sharding.shard_map(sm, edges, r))
    assert nq == cs.shape[0] + ss.shape[0]  # every query has exactly one owner
    errors = []

    def run(r):
        try:
            for it in range(3):
                got[r][it] = ranks[r].map_iter_allreduce(it, T)
        except BaseException as e:
            errors.append(e)

    th = [threading.Thread(target=run, args=(r,)) for r in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join(60)
    assert not errors, errors
    for it in range(3):
        (a0, b0, n0), (a1, b1, n1) = got[0][it], got[1][it]
        assert n0 == n1 and np.array_equal(a0, a1) and np.array_equal(b0, b1)  # every rank holds the same global sums
        wa, wb, wn = want[it]
        assert n0 == wn and wn > 1000
        # the shards add the same exact products in another order: equal up to the single fp32 rounding at the end
        assert np.abs(a0 - wa).max() <= 1e-6 * np.abs(wa).max() and np.abs(b0 - wb).max() <= 1e-6 * np.abs(wb).max()
    for g in ranks + [whole]:
        g.close()
