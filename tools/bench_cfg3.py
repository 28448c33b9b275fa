"""cfg 3 (BASELINE.json configs[2]): HDL-64-shaped sweeps (64 x 1875 = 120 k rays) registered scan-to-map against a
pre-filled ~2 M-point local map on one B200.

The map is pre-filled through the library itself: the scene is mapped once with the full pipeline (so cubes hold real
voxel-gridded clouds), then the same cubes are densified by registering jittered copies until the gathered local map
reaches the requested size.  Measured: sweeps/s of the full pipeline at that map size, per-kernel-class time and
achieved algorithmic GB/s vs the measured HBM peak.
"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


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
