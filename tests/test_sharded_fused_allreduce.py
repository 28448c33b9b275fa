"""SURVEY 8e (2), "optimised" variant: the 28-double all-reduce fused into the Gauss-Newton kernel (peer stores + flags).

Two shards of one map, two handles ("ranks") on ONE GPU.  Kernels of different ranks must never wait for one another on
one GPU (nothing guarantees they run at the same time), so the ranks run one after the other: each rank's partial sums are
computed first (loam_map_iter_partial), deposited in the OTHER rank's exchange buffer exactly as that peer's kernel would
store them over NVLink (loam_shard_inject), and then the rank's fused iteration runs -- store to the peer, find the peer's
flag already raised, add the partials in rank order.  The global sums must equal the un-sharded iteration's.  Every rank
holds the whole stack and picks its own queries on the device (loam_shard_set_slab).  The real multi-GPU exchange (CUDA
IPC, several iterations inside one launch) is covered by tools/bench_cfg5.py under `gpurun --gpus N`.
"""
import numpy as np
import pytest


@pytest.mark.gpu
def test_fused_allreduce_two_ranks_one_gpu(orc, sweeps16):
    import torch
    from gpscalibration_b200 import LoamGpu, sharding
    from helpers.routing import owner_mask
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
    own = [int(owner_mask(orc, cs, T, edges, r).sum() + owner_mask(orc, ss, T, edges, r).sum()) for r in range(2)]
    assert sum(own) == cs.shape[0] + ss.shape[0] and min(own) > 100  # every query has exactly one owner; the split is real
    ranks = [LoamGpu(), LoamGpu()]
    handles = [g.shard_export() for g in ranks]
    for r, g in enumerate(ranks):
        g.shard_connect(handles, r)
        g.shard_set_slab(edges[r], edges[r + 1])
        g.map_set_inputs(cs, ss, sharding.shard_map(cm, edges, r), sharding.shard_map(sm, edges, r))
    part = torch.zeros(32, dtype=torch.float64, device="cuda")
    for it in range(3):
        partial = []
        for g in ranks:
            g.map_iter_partial(it, T, part.data_ptr())
            partial.append(part[:28].cpu().numpy().copy())
        assert int(partial[0][27]) + int(partial[1][27]) == want[it][2]
        got = []
        for r, g in enumerate(ranks):
            g.shard_inject(1 - r, partial[1 - r])
            got.append(g.map_iter_allreduce(it, T))
        (a0, b0, n0), (a1, b1, n1) = got
        assert n0 == n1 and np.array_equal(a0, a1) and np.array_equal(b0, b1)  # every rank holds the same global sums
        wa, wb, wn = want[it]
        assert n0 == wn and wn > 1000
        # the shards add the same exact products in another order: equal up to the single fp32 rounding at the end
        assert np.abs(a0 - wa).max() <= 1e-6 * np.abs(wa).max() and np.abs(b0 - wb).max() <= 1e-6 * np.abs(wb).max()
    for g in ranks + [whole]:
        g.close()
