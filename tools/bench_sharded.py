"""cfg 5 (BASELINE.json configs[4]): one large query cloud against a large map sharded in x-slabs over the ranks,
28-double all-reduce per Gauss-Newton iteration (NCCL), identical normal equations on every rank.

  torchrun --nproc-per-node N tools/bench_sharded.py --map-points 20000000 --queries 1000000
Prints one JSON line on rank 0: iterations/s, per-iteration ms (max over ranks), all-reduce us, per-GPU algorithmic
GB/s of the kNN+fit pair (96 B per query) and of the index build (36 B per map point).
"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def synth_map(n_surf, n_corner, extent, seed=7):
    """Planar 'city': ground plane + wall planes every 40 m (surf) and vertical edges (corner), ~0.4 / 0.2 m spacing jitter."""
    rng = np.random.default_rng(seed)
    ng = int(n_surf * 0.6)
    g = np.empty((ng, 4), np.float32)
    g[:, 0] = rng.uniform(-extent, extent, ng); g[:, 2] = rng.uniform(-extent / 2, extent / 2, ng)
    g[:, 1] = -1.8 + rng.normal(0, 0.01, ng); g[:, 3] = 0
    nw = n_surf - ng
    w = np.empty((nw, 4), np.float32)
    w[:, 0] = rng.uniform(-extent, extent, nw); w[:, 1] = rng.uniform(-1.8, 10, nw)
    w[:, 2] = (rng.integers(-6, 7, nw) * 40.0 + rng.normal(0, 0.01, nw)).astype(np.float32); w[:, 3] = 0
    c = np.empty((n_corner, 4), np.float32)
    c[:, 0] = rng.integers(-int(extent / 10), int(extent / 10) + 1, n_corner) * 10.0 + rng.normal(0, 0.01, n_corner)
    c[:, 2] = rng.integers(-6, 7, n_corner) * 40.0 + rng.normal(0, 0.01, n_corner)
    c[:, 1] = rng.uniform(-1.8, 10, n_corner); c[:, 3] = 0
    return c, np.concatenate([g, w])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--map-points", type=int, default=10_000_000)
    ap.add_argument("--queries", type=int, default=1_000_000)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--extent", type=float, default=2000.0)
    args = ap.parse_args()
    import torch, torch.distributed as dist
    from gpscalibration_b200 import LoamGpu, capi, sharding
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_corner, n_surf = args.map_points // 5, args.map_points - args.map_points // 5
    corner_map, surf_map = synth_map(n_surf, n_corner, args.extent)
    rng = np.random.default_rng(11)
    T_true = np.array([0.002, 0.01, -0.003, 0.05, -0.02, 0.08], np.float32)
    qi_s = rng.choice(n_surf, args.queries * 4 // 5, replace=False); qi_c = rng.choice(n_corner, args.queries // 5, replace=False)
    # queries = map points pushed through the INVERSE of T_true (approximately: small angles), so T converges towards T_true
    def inv(p):
        q = p.copy(); q[:, :3] -= T_true[3:]
        return q
    surf_stack, corner_stack = inv(surf_map[qi_s]), inv(corner_map[qi_c])
    edges = sharding.slab_edges(-args.extent, args.extent, world)
    T = np.zeros(6, np.float32)
    t0 = time.time()
    my_cm, my_sm = sharding.shard_map(corner_map, edges, rank), sharding.shard_map(surf_map, edges, rank)
    my_cs, my_ss = sharding.route_queries(corner_stack, T, edges, rank), sharding.route_queries(surf_stack, T, edges, rank)
    gpu = LoamGpu(device=local)
    stream = torch.cuda.ExternalStream(gpu.stream, device=torch.device("cuda", local))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(stream)
    gpu.map_set_inputs(my_cs, my_ss, my_cm, my_sm)  # upload + voxel-hash build
    e1.record(stream); torch.cuda.synchronize()
    build_ms = e0.elapsed_time(e1)
    part = torch.zeros(32, dtype=torch.float64, device="cuda")
    state = np.zeros(37, np.float32)
    it_ms, ar_us = [], []
    gpu.profile(True)
    for it in range(args.iters):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(); w0 = time.perf_counter()
        gpu.map_iter_partial(it, T, part.data_ptr())
        torch.cuda.synchronize(); w1 = time.perf_counter()
        if world > 1:
            dist.all_reduce(part, op=dist.ReduceOp.SUM)
        red = part[:28].cpu().numpy(); w2 = time.perf_counter()
        AtA, AtB, n_sel = capi.finish_reduced(red)
        X = capi.gn_solve(AtA, AtB, it, 100.0, state)
        T = (T + X).astype(np.float32)
        it_ms.append(1e3 * (w2 - w0)); ar_us.append(1e6 * (w2 - w1))
    prof = gpu.profile_read()
    gpu.profile(False)
    t_it = torch.tensor([np.median(it_ms[2:])], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_it, op=dist.ReduceOp.MAX)
    nq = my_cs.shape[0] + my_ss.shape[0]
    kern_ms = np.median(it_ms[2:]) - np.median(ar_us[2:]) / 1e3
    out = {"config": "cfg5 sharded map", "n_gpus": world, "map_points": args.map_points, "queries": args.queries, "rank0_map_points": int(my_cm.shape[0] + my_sm.shape[0]),
           "rank0_queries": int(nq), "n_sel": int(n_sel), "iter_ms_max_over_ranks": float(t_it.item()), "iterations_per_s": 1e3 / float(t_it.item()),
           "allreduce_plus_d2h_us": float(np.median(ar_us[2:])), "index_build_ms": build_ms,
           "knn_fit_GBps_algorithmic": 96.0 * nq / (kern_ms * 1e-3) / 1e9, "index_build_GBps_algorithmic_incl_h2d": 36.0 * (my_cm.shape[0] + my_sm.shape[0]) / (build_ms * 1e-3) / 1e9,
           "kernel_ms_per_iter": {k: round(prof[k]["ms"] / args.iters, 4) for k in ("map_knn", "map_fit")},
           "knn_fit_GBps_kernel_only": 96.0 * nq / ((prof["map_knn"]["ms"] + prof["map_fit"]["ms"]) / args.iters * 1e-3) / 1e9,
           "T_final": [round(float(x), 5) for x in T], "T_true": [float(x) for x in T_true]}
    if rank == 0:
        print(json.dumps(out))
    gpu.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
