"""world_size-2 gloo test (CPU) of the sharded-map host logic: slab + halo sharding of the local map, routing of the
queries, 28-double all-reduce, loam_map_finish_reduced — must give the normal equations of the unsharded iteration."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from gpscalibration_b200 import capi, sharding, SweepGenerator
    from oracle import orc
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers.routing import owner_mask
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # every rank rebuilds the same deterministic inputs: a local map from a few sweeps and the next sweep's stacks
    gen = SweepGenerator()
    pipe = orc.Pipeline()
    for k in range(8):
        pipe.process(gen.sweep(k)[0])
    corner_map = np.concatenate([pipe.cloud("corner_last")])  # registered features of the last sweeps stand in for a map
    surf_map = pipe.cloud("registered")[::3].copy()
    sr = orc.ScanRegistration()
    f = sr.extract(gen.sweep(8)[0])
    corner_stack, surf_stack = orc.voxel_grid(f["less_sharp"], 0.2), orc.voxel_grid(f["less_flat"], 0.4)
    T = np.array(pipe.process(gen.sweep(8)[0]).mapped, np.float32)
    # the map clouds above live in different frames than a real local map; what matters here is only that sharded ==
    # unsharded on identical inputs, so bring the map into the query frame with the same pose
    surf_map = orc.associate_to_map(orc.voxel_grid(f["less_flat"], 0.2), T)
    corner_map = orc.associate_to_map(orc.voxel_grid(f["less_sharp"], 0.1), T)
    edges = sharding.slab_edges(float(surf_map[:, 0].min()), float(surf_map[:, 0].max()), world)
    my = orc.map_iteration_sums28(corner_stack[owner_mask(orc, corner_stack, T, edges, rank)],
                                  surf_stack[owner_mask(orc, surf_stack, T, edges, rank)],
                                  sharding.shard_map(corner_map, edges, rank), sharding.shard_map(surf_map, edges, rank), T)
    t = torch.from_numpy(my.copy())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    full = orc.map_iteration_sums28(corner_stack, surf_stack, corner_map, surf_map, T)
    AtA, AtB, n = capi.finish_reduced(t.numpy())
    rAtA, rAtB, rn = capi.finish_reduced(full)
    q.put((rank, int(my[27]), n, rn, float(np.abs(AtA - rAtA).max() / max(1e-30, np.abs(rAtA).max())),
           float(np.abs(t.numpy() - full).max() / np.abs(full).max())))
    dist.destroy_process_group()


def test_sharded_normal_equations_equal_unsharded():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=240) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    shares = sorted(r[1] for r in res)
    assert shares[0] > 0, "both ranks must own rows"  # the split is real
    for rank, mine, n, rn, rel_f32, rel_f64 in res:
        assert n == rn and n == sum(shares) and n > 200
        assert rel_f32 == 0.0          # rounded normal equations identical to the unsharded iteration
        assert rel_f64 < 1e-12         # double sums equal up to summation order


def test_slab_sharding_covers_every_accepted_neighbour():
    sys.path.insert(0, ROOT)
    from gpscalibration_b200 import sharding
    rng = np.random.default_rng(0)
    cloud = np.zeros((20000, 4), np.float32)
    cloud[:, :3] = rng.uniform(-50, 50, (20000, 3))
    edges = sharding.slab_edges(-50, 50, 4)
    q = rng.uniform(-49, 49, (500, 3)).astype(np.float32)
    for r in range(4):
        own = q[(q[:, 0] >= edges[r]) & (q[:, 0] < edges[r + 1])]
        shard = sharding.shard_map(cloud, edges, r)
        for p in own:
            d2_all = ((cloud[:, :3] - p) ** 2).sum(1)
            d2_sh = ((shard[:, :3] - p) ** 2).sum(1)
            assert (d2_all < 1.0).sum() == (d2_sh < 1.0).sum()  # nothing within the 1 m gate is lost
