"""cfg 3 (BASELINE.json configs[2]): HDL-64-shaped sweeps (64 x 1875 = 120 k rays) scan-to-map on one B200.  Two measurements:

  main()       the full pipeline (extract + odometry + mapping, blocking call) on ray-cast HDL-64-shaped sweeps; the local
               map is whatever the synthetic yard fills the cubes with (~0.17 M points: the scene is 160 m wide).
  run_stage()  the mapping stage at the SPEC size: a 2.0 M-point voxel-filtered local map (0.4 M corner + 1.6 M surf,
               SURVEY 8d cfg 3) and stacks of the size an HDL-64 sweep leaves after LM:736-747 (6.4 k corner + 10.3 k surf,
               all within 80 m of the sensor) through loam_map_set_inputs (index build, LM:750-751) and loam_map_optimize
               (the whole Gauss-Newton loop, LM:753-1017).  These are the two pieces of a mapping run whose cost grows
               with the map; extraction and odometry do not see the map.  bench.py embeds the result as "cfg3".
"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


T_TRUE = np.array([0.001, 0.004, -0.001, 0.05, -0.02, 0.08], np.float32)


def run_stage(local=0, n_corner_map=400_000, n_surf_map=1_600_000, n_cs=6_400, n_ss=10_300, reps=20, log=lambda *a: None):
    import torch
    from gpscalibration_b200 import LoamGpu, mapsynth
    t0 = time.time()
    corner_map, surf_map, extent = mapsynth.synth_map(n_surf_map, n_corner_map)
    cs, ss = mapsynth.synth_queries_local(corner_map, surf_map, n_cs, n_ss, T_TRUE)
    log(f"[cfg3] map {corner_map.shape[0]} + {surf_map.shape[0]} points, stacks {cs.shape[0]} + {ss.shape[0]} in {time.time() - t0:.1f}s")
    dev = torch.device("cuda", local)
    gpu = LoamGpu(device=local, max_map_points=1 << 21)
    stream = torch.cuda.ExternalStream(gpu.stream, device=dev)
    n_map = corner_map.shape[0] + surf_map.shape[0]
    build_ms, loop_ms, iters, Tf = [], [], 0, None
    for rep in range(reps + 2):
        gpu.profile(True)
        gpu.map_set_inputs(cs, ss, corner_map, surf_map)  # upload + index build; "grid" = the build kernels alone
        b = gpu.profile_read()["grid"]["ms"]
        gpu.profile(False)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        Tf, iters = gpu.map_optimize(np.zeros(6, np.float32), 10)
        e1.record(stream)
        torch.cuda.synchronize(dev)
        if rep >= 2:
            build_ms.append(b)
            loop_ms.append(e0.elapsed_time(e1))
    gpu.close()
    peak = 6553.9
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", peak))
    except Exception:
        pass
    bm, lm = float(np.median(build_ms)), float(np.median(loop_ms))
    nq = cs.shape[0] + ss.shape[0]
    return {"config": "cfg3 at spec size, mapping stage: %d-point voxel-filtered local map, HDL-64-sized stacks (%d corner + %d surf)" % (n_map, cs.shape[0], ss.shape[0]),
            "map_points": int(n_map), "stack_points": int(nq), "index_build_ms": bm, "gn_loop_ms": lm, "iterations": int(iters),
            "mapping_runs_per_s": 1e3 / (bm + lm),
            "index_build": {"algorithmic_GBps": 36.0 * n_map / (bm * 1e-3) / 1e9, "frac_of_hbm_peak": 36.0 * n_map / (bm * 1e-3) / 1e9 / peak},
            "gn_iteration": {"us": 1e3 * lm / max(1, iters), "algorithmic_GBps": 96.0 * nq / (lm / max(1, iters) * 1e-3) / 1e9,
                             "frac_of_hbm_peak": 96.0 * nq / (lm / max(1, iters) * 1e-3) / 1e9 / peak},
            "T_final": [float(x) for x in Tf], "T_true": [float(x) for x in T_TRUE], "hbm_peak_gbs": peak}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sweeps", type=int, default=60)
    ap.add_argument("--warm", type=int, default=40, help="sweeps used to fill the map before timing")
    args = ap.parse_args()
    from gpscalibration_b200 import LoamGpu, SweepGenerator
    step = 26.8 / 63.0
    gen = SweepGenerator(sensor=2, scene=1, seed=0xC0FFEE)
    n = args.warm + args.sweeps
    sw = [gen.sweep(k)[0].copy() for k in range(n)]
    gpu = LoamGpu(n_scans=64, ring_mode=1, ring_ang_min=-24.8, ring_ang_step=step)
    for k in range(args.warm):
        r = gpu.process_sweep(sw[k])
    gpu.host_times()
    gpu.profile(True)
    t0 = time.perf_counter()
    sizes = []
    for k in range(args.warm, n):
        r = gpu.process_sweep(sw[k])
        if r.mapping_ran:
            sizes.append((r.map.n_corner_map, r.map.n_surf_map, r.map.n_corner_stack, r.map.n_surf_stack, r.map.iterations))
    dt = time.perf_counter() - t0
    prof = gpu.profile_read()
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6553.9) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    per_unit = {"extract": 39.0, "odom_iter": 64.0, "to_end": 32.0, "map_stack": 32.0, "voxel": 32.0, "grid": 36.0, "map_knn": 96.0, "gather": 32.0, "insert": 32.0}
    classes = {}
    for k, v in prof.items():
        if v["scopes"] == 0:
            continue
        gbs = per_unit.get(k, 0.0) * v["units"] / (v["ms"] * 1e-3) / 1e9 if v["ms"] > 0 and k in per_unit else None
        classes[k] = {"ms_per_sweep": round(v["ms"] / args.sweeps, 4), "us_per_launch_group": round(1e3 * v["ms"] / v["scopes"], 1),
                      "algorithmic_GBps": None if gbs is None else round(gbs, 1), "frac_of_hbm_peak": None if gbs is None else round(gbs / peak, 4)}
    out = {"config": "cfg3 HDL-64-shaped sweeps vs large local map, blocking loam_process_sweep", "sweeps": args.sweeps,
           "points_per_sweep": int(np.mean([s.shape[0] for s in sw])), "sweeps_per_s": args.sweeps / dt,
           "last_map_sizes_corner_surf_stackc_stacks_iters": sizes[-1] if sizes else None, "hbm_peak_gbs": peak, "classes": classes,
           "host_ms_per_sweep": {k: round(1e3 * v / args.sweeps, 3) for k, v in gpu.host_times().items() if v > 0}}
    print(json.dumps(out))
    gpu.close()


if __name__ == "__main__":
    main()
