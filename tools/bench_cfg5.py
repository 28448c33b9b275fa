"""cfg 5 (BASELINE.json configs[4]): one 1 M-point query stack against a 50 M-point voxel-filtered map, sharded in x-slabs
(+ 1 m halo) over the ranks; per Gauss-Newton iteration the ranks exchange 28 doubles.

  python tools/bench_cfg5.py --map-points 50000000 --queries 1000000                       # one GPU
  torchrun --nproc-per-node N tools/bench_cfg5.py --map-points 50000000 --queries 1000000  # N = 2, 4, 8

Every rank holds the WHOLE stack and its slab of the map; which stack points a rank evaluates is decided on the device
every iteration (loam_shard_set_slab).  Two implementations of the exchange are timed on the same inputs:
  nccl   loam_map_iter_partial -> NCCL all-reduce (torch.distributed) -> D2H -> host solve: one launch + one collective
         + one host round trip per iteration
  fused  loam_map_optimize: ONE launch for the whole loop; the last CTA of every rank stores its 28 sums into the peers'
         exchange buffers over NVLink (CUDA IPC), waits for theirs, solves and goes on -- the host sees the final pose
Prints one JSON line on rank 0 (`run()` returns the dict: bench.py embeds it in its own line under "cfg5").
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# The stack is the map moved by the inverse of this pose.  The map is 3.2 km wide: rotations of 1e-5 rad move its far end
# by 1.6 cm -- the size of error scan-to-map registration corrects (it starts from the odometry pose, LM:465) -- so the
# Gauss-Newton loop converges like it does in the pipeline.  (Round 1 used 1e-2 rad: 16 m at the far end, where a point's
# nearest neighbours are no longer its own surface and the loop just runs into the 10-iteration limit.)
T_TRUE = np.array([2e-6, 1e-5, -3e-6, 0.05, -0.02, 0.08], np.float32)


def run(map_points=50_000_000, queries=1_000_000, iters=10, rank=0, local=0, world=1, dist=None, unfiltered=False,
        random_order=False, fused=True, log=lambda *a: None, t_true=None):
    import torch
    from gpscalibration_b200 import LoamGpu, capi, mapsynth, sharding
    dev = torch.device("cuda", local)
    T_TRUE = np.asarray(t_true, np.float32) if t_true is not None else globals()["T_TRUE"]
    t0 = time.time()
    if unfiltered:
        corner_map, surf_map, extent = mapsynth.synth_map_unfiltered(map_points - map_points // 5, map_points // 5, 2000.0)
    else:
        corner_map, surf_map, extent = mapsynth.synth_map(map_points - map_points // 5, map_points // 5)
    cs, ss = mapsynth.synth_queries(corner_map, surf_map, queries, T_TRUE, ordered=not random_order)
    n_map = corner_map.shape[0] + surf_map.shape[0]
    edges = sharding.slab_edges(-extent, extent, world)
    my_cm, my_sm = sharding.shard_map(corner_map, edges, rank), sharding.shard_map(surf_map, edges, rank)
    del corner_map, surf_map
    log(f"[cfg5 rank {rank}] map {n_map} points generated + sharded in {time.time() - t0:.1f}s; mine {my_cm.shape[0] + my_sm.shape[0]}")
    my_map = my_cm.shape[0] + my_sm.shape[0]
    gpu = LoamGpu(device=local)  # the stage-level calls size their buffers on demand
    if world > 1:
        gpu.shard_set_slab(edges[rank], edges[rank + 1])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    gpu.profile(True)
    gpu.map_set_inputs(cs, ss, my_cm, my_sm)  # upload + index build; the build alone is timed by the profile class "grid"
    build_ms = gpu.profile_read()["grid"]["ms"]
    gpu.profile(False)
    nq = cs.shape[0] + ss.shape[0]

    # ---- baseline: one launch + NCCL all-reduce + host solve per iteration
    part = torch.zeros(32, dtype=torch.float64, device=dev)
    T = np.zeros(6, np.float32)
    state = np.zeros(37, np.float32)
    it_ms, ar_us, kern_ms, n_sel = [], [], [], 0
    iters_nccl = 0
    for it in range(iters):
        barrier()
        w0 = time.perf_counter()
        gpu.map_iter_partial(it, T, part.data_ptr())
        w1 = time.perf_counter()
        if world > 1:
            dist.all_reduce(part, op=dist.ReduceOp.SUM)
        red = part[:28].cpu().numpy()
        w2 = time.perf_counter()
        AtA, AtB, n_sel = capi.finish_reduced(red)
        iters_nccl = it + 1
        it_ms.append(1e3 * (w2 - w0)); ar_us.append(1e6 * (w2 - w1)); kern_ms.append(1e3 * (w1 - w0))
        if n_sel < 50:
            continue
        X = capi.gn_solve(AtA, AtB, it, 100.0, state)
        T = (T + X).astype(np.float32)
        dR = np.float32(np.sqrt(((X[:3].astype(np.float64) * 180.0 / np.pi) ** 2).sum()))
        dT = np.float32(np.sqrt(((X[3:].astype(np.float64) * 100) ** 2).sum()))
        if dR < 0.05 and dT < 0.05:
            break
    own = int(red[27]) if world == 1 else None
    nccl_iter_ms = max_over_ranks(np.median(it_ms[1:]) if len(it_ms) > 1 else it_ms[0])
    kernel_iter_ms = max_over_ranks(np.median(kern_ms[1:]) if len(kern_ms) > 1 else kern_ms[0])

    out = {"config": "cfg5: %d-point voxel-filtered map sharded in x-slabs (+1 m halo), %d-point stack, %d GPU(s)" % (n_map, nq, world),
           "n_gpus": world, "map_points": int(n_map), "queries": int(nq), "rank0_map_points": int(my_map),
           "query_order": "random" if random_order else "voxel-grid (ascending cell id)", "n_sel": int(n_sel),
           "index_build_ms_kernels_only": build_ms, "index_build_GBps_algorithmic": 36.0 * my_map / (build_ms * 1e-3) / 1e9 if build_ms > 0 else None,
           "nccl": {"iterations": iters_nccl, "iter_ms_max_over_ranks": nccl_iter_ms, "kernel_ms_max_over_ranks": kernel_iter_ms,
                    "allreduce_plus_d2h_us": float(np.median(ar_us[1:]) if len(ar_us) > 1 else ar_us[0]),
                    "T_final": [float(x) for x in T]},
           "T_true": [float(x) for x in T_TRUE]}

    # ---- fused: the whole loop in one launch, exchange inside the kernel
    if fused:
        if world > 1:
            hd = gpu.shard_export()
            mine = torch.frombuffer(bytearray(hd), dtype=torch.uint8).to(dev)
            gathered = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(gathered, mine)
            gpu.shard_connect([bytes(g.cpu().numpy().tobytes()) for g in gathered], rank)
        stream = torch.cuda.ExternalStream(gpu.stream, device=dev)
        f_ms, Tf, itf = [], None, 0
        for rep in range(4):
            gpu.map_set_inputs(cs, ss, my_cm, my_sm)  # fresh matP / isDegenerate state, same index
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            w0 = time.perf_counter()
            Tf, itf = gpu.map_optimize(np.zeros(6, np.float32), iters)
            w1 = time.perf_counter()
            e1.record(stream)
            torch.cuda.synchronize(dev)
            f_ms.append((e0.elapsed_time(e1), 1e3 * (w1 - w0)))
        dev_ms = max_over_ranks(float(np.median([a for a, _ in f_ms[1:]])))
        wall_ms = max_over_ranks(float(np.median([b for _, b in f_ms[1:]])))
        same = bool(np.array_equal(Tf, T)) and itf == iters_nccl
        if world > 1:  # every rank must hold the same pose, bit for bit
            tt = torch.from_numpy(Tf.copy()).to(dev)
            allT = [torch.empty_like(tt) for _ in range(world)]
            dist.all_gather(allT, tt)
            same = same and all(bool(torch.equal(allT[0], a)) for a in allT)
        out["fused"] = {"iterations": int(itf), "loop_ms_device_max_over_ranks": dev_ms, "loop_ms_wall_max_over_ranks": wall_ms,
                        "iter_ms": dev_ms / max(1, itf), "iterations_per_s": 1e3 * itf / dev_ms,
                        "T_final_equals_nccl_path_on_every_rank": same, "T_final": [float(x) for x in Tf],
                        "per_gpu_GBps_algorithmic": 96.0 * (nq / world) / (dev_ms / max(1, itf) * 1e-3) / 1e9}
        out["iterations_per_s"] = out["fused"]["iterations_per_s"]
    gpu.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--map-points", type=int, default=50_000_000)
    ap.add_argument("--queries", type=int, default=1_000_000)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--unfiltered-map", action="store_true", help="worst case: map NOT voxel-filtered (crowded cells)")
    ap.add_argument("--random-query-order", action="store_true", help="worst case: no spatial coherence between consecutive queries")
    ap.add_argument("--no-fused", action="store_true")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    out = run(args.map_points, args.queries, args.iters, rank, local, world, dist if world > 1 else None, args.unfiltered_map,
              args.random_query_order, not args.no_fused, log=lambda *a: print(*a, file=sys.stderr, flush=True))
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
