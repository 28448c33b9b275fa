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


def synth_map_unfiltered(n_surf, n_corner, extent, seed=7):
    """Worst case (--unfiltered-map): uniformly random points on the planes / edges, NOT voxel-filtered, so wall cells hold
    ~11 and edge cells ~30 points (a real LM map never does: every cube is re-filtered at 0.2 / 0.4 m, LM:642-662)."""
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
    return c, np.concatenate([g, w]), extent


def synth_map(n_surf, n_corner, seed=7):
    """Planar 'city' as LM keeps it: every surface carries one point per 0.4 m voxel (surf) and every vertical edge one
    point per 0.2 m voxel (corner), i.e. what the per-cube voxel filter (LM:642-662) leaves.  Ground plane 2E x E,
    13 wall planes across it, edges on the walls; E follows from the requested point counts.  Returns (corner, surf, E)."""
    rng = np.random.default_rng(seed)
    n_g = int(n_surf * 0.8)
    nz = int(np.sqrt(n_g / 2)); nx = 2 * nz
    E = 0.4 * nz
    gx, gz = np.meshgrid(np.arange(nx, dtype=np.float32), np.arange(nz, dtype=np.float32), indexing="ij")
    g = np.empty((nx * nz, 4), np.float32)
    g[:, 0] = (gx.ravel() + rng.uniform(0.1, 0.9, nx * nz)) * 0.4 - E
    g[:, 2] = (gz.ravel() + rng.uniform(0.1, 0.9, nx * nz)) * 0.4 - E / 2
    g[:, 1] = -1.8 + rng.normal(0, 0.01, nx * nz); g[:, 3] = 0
    n_w = n_surf - nx * nz
    ny = max(1, n_w // (13 * nx))
    wx, wy, wk = np.meshgrid(np.arange(nx, dtype=np.float32), np.arange(ny, dtype=np.float32), np.arange(13, dtype=np.float32), indexing="ij")
    m = nx * ny * 13
    w = np.empty((m, 4), np.float32)
    w[:, 0] = (wx.ravel() + rng.uniform(0.1, 0.9, m)) * 0.4 - E
    w[:, 1] = (wy.ravel() + rng.uniform(0.1, 0.9, m)) * 0.4 - 1.8
    w[:, 2] = (wk.ravel() - 6) * (E / 13.0) + rng.normal(0, 0.01, m); w[:, 3] = 0
    H = ny * 0.4
    nyc = max(1, int(H / 0.2))
    nl = max(1, n_corner // (13 * nyc))  # edges per wall
    cx, cy, ck = np.meshgrid(np.arange(nl, dtype=np.float32), np.arange(nyc, dtype=np.float32), np.arange(13, dtype=np.float32), indexing="ij")
    m = nl * nyc * 13
    c = np.empty((m, 4), np.float32)
    c[:, 0] = (cx.ravel() + 0.5) * (2 * E / nl) - E + rng.normal(0, 0.01, m)
    c[:, 1] = (cy.ravel() + rng.uniform(0.1, 0.9, m)) * 0.2 - 1.8
    c[:, 2] = (ck.ravel() - 6) * (E / 13.0) + rng.normal(0, 0.01, m); c[:, 3] = 0
    return c, np.concatenate([g, w]), E


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--map-points", type=int, default=10_000_000)
    ap.add_argument("--queries", type=int, default=1_000_000)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--extent", type=float, default=2000.0, help="half-length in x of the --unfiltered-map scene")
    ap.add_argument("--fused-allreduce", action="store_true", help="also run loam_map_iter_allreduce (P2P one-shot all-reduce in the kernel)")
    ap.add_argument("--unfiltered-map", action="store_true", help="worst case: map NOT voxel-filtered (crowded cells)")
    ap.add_argument("--random-query-order", action="store_true", help="worst case: no spatial coherence between consecutive queries")
    args = ap.parse_args()
    import torch, torch.distributed as dist
    from gpscalibration_b200 import LoamGpu, capi, sharding
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if args.unfiltered_map:
        n_corner, n_surf = args.map_points // 5, args.map_points - args.map_points // 5
        corner_map, surf_map, extent = synth_map_unfiltered(n_surf, n_corner, args.extent)
    else:
        n_corner, n_surf = args.map_points // 20, args.map_points - args.map_points // 20
        corner_map, surf_map, extent = synth_map(n_surf, n_corner)
    n_corner, n_surf = len(corner_map), len(surf_map)
    rng = np.random.default_rng(11)
    T_true = np.array([0.002, 0.01, -0.003, 0.05, -0.02, 0.08], np.float32)
    qi_c = rng.choice(n_corner, min(args.queries // 5, n_corner // 2), replace=False); qi_s = rng.choice(n_surf, args.queries - len(qi_c), replace=False)
    # queries = map points pushed through the INVERSE of T_true (approximately: small angles), so T converges towards T_true
    def inv(p):
        q = p.copy(); q[:, :3] -= T_true[3:]
        return q
    surf_stack, corner_stack = inv(surf_map[qi_s]), inv(corner_map[qi_c])

    def voxel_order(c, leaf):  # stack clouds are voxel-grid outputs: ascending (k, j, i) cell order (PCL's linear id)
        ijk = np.floor(c[:, :3] / leaf).astype(np.int64)
        return c[np.lexsort((ijk[:, 0], ijk[:, 1], ijk[:, 2]))]
    if not args.random_query_order:
        surf_stack, corner_stack = voxel_order(surf_stack, 0.4), voxel_order(corner_stack, 0.2)
    edges = sharding.slab_edges(-extent, extent, world)
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
    # ---- the same iterations with the all-reduce fused into the reduction kernel (NVLink peer stores, CUDA IPC)
    fused = None
    if args.fused_allreduce:
        hd = gpu.shard_export()
        if world > 1:
            mine = torch.frombuffer(bytearray(hd), dtype=torch.uint8).cuda()
            gathered = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(gathered, mine)
            handles = [bytes(g.cpu().numpy().tobytes()) for g in gathered]
        else:
            handles = [hd]
        gpu.shard_connect(handles, rank)
        T2, state2, f_ms = np.zeros(6, np.float32), np.zeros(37, np.float32), []
        for it in range(args.iters):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize(); w0 = time.perf_counter()
            AtA, AtB, n2 = gpu.map_iter_allreduce(it, T2)
            w1 = time.perf_counter()
            X = capi.gn_solve(AtA, AtB, it, 100.0, state2)
            T2 = (T2 + X).astype(np.float32)
            f_ms.append(1e3 * (w1 - w0))
        tf = torch.tensor([np.median(f_ms[2:])], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tf, op=dist.ReduceOp.MAX)
        fused = {"iter_ms_max_over_ranks": float(tf.item()), "n_sel": int(n2), "T_final_equals_nccl_path": bool(np.array_equal(T2, T)),
                 "T_final": [float(x) for x in T2]}
    t_it = torch.tensor([np.median(it_ms[2:])], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_it, op=dist.ReduceOp.MAX)
    nq = my_cs.shape[0] + my_ss.shape[0]
    kern_ms = np.median(it_ms[2:]) - np.median(ar_us[2:]) / 1e3
    out = {"config": "cfg5 sharded map", "n_gpus": world, "map_points": args.map_points, "queries": args.queries, "rank0_map_points": int(my_cm.shape[0] + my_sm.shape[0]),
           "rank0_queries": int(nq), "query_order": "random" if args.random_query_order else "voxel-grid (ascending cell id)", "n_sel": int(n_sel), "iter_ms_max_over_ranks": float(t_it.item()), "iterations_per_s": 1e3 / float(t_it.item()),
           "allreduce_plus_d2h_us": float(np.median(ar_us[2:])), "index_build_ms": build_ms,
           "knn_fit_GBps_algorithmic": 96.0 * nq / (kern_ms * 1e-3) / 1e9, "index_build_GBps_algorithmic_incl_h2d": 36.0 * (my_cm.shape[0] + my_sm.shape[0]) / (build_ms * 1e-3) / 1e9,
           "kernel_ms_per_iter": {k: round(prof[k]["ms"] / args.iters, 4) for k in ("map_knn", "map_fit")},
           "knn_fit_GBps_kernel_only": 96.0 * nq / ((prof["map_knn"]["ms"] + prof["map_fit"]["ms"]) / args.iters * 1e-3) / 1e9,
           "T_final": [round(float(x), 5) for x in T], "T_true": [float(x) for x in T_true]}
    if fused is not None:
        out["fused_allreduce"] = fused
    if rank == 0:
        print(json.dumps(out))
    gpu.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
