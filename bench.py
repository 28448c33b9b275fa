#!/usr/bin/env python
"""bench.py — LiDAR sweeps/s through extract + scan-to-scan odometry + scan-to-map mapping (BASELINE.json metric).

Workload (BASELINE.json configs[1]): a synthetic 1000-sweep VLP-16-shaped sequence (16 rings x 1800 columns, the
reference's own ring-table angles, seed 0xC0FFEE + 1000 * rank) through the full hot path on one B200 per rank.
One STEP = reset + one pass over the whole sequence.  With N > 1 ranks every rank registers its own independent
sequence (BASELINE configs[3], no data-path collective): weak scaling, value = all sweeps / max-over-ranks time.

  value : device-timed (CUDA events bracketing all stage streams), the sweeps already resident in HBM, pipelined C ABI
  e2e   : the same through the reference-facing C-ABI call with HOST buffers (pinned): H2D of every sweep and the
          D2H reads of counts / 28-double normal-equation mailboxes / poses inside the timed region
  sync_api : both numbers again through the blocking one-call-per-sweep entry point (loam_process_sweep)
  roofline : dominant kernel class from a separate CUDA-event pass (loam_profile), algorithmic bytes per DESIGN.md
  cpu_baseline : the CPU oracle (restatement of the reference; oracle/_ref when built) on this box's host cores

Extra objects on the same line: "cfg1" (configs[0]: microseconds per odometry refresh / iteration of one registration
against the launch-latency floor), "cfg3" (configs[2] at spec size: mapping stage against a 2.0 M-point local map, N = 1),
"cfg5" (configs[4]: 1 M-point stack against a 50 M-point map sharded over the N ranks, NCCL all-reduce and the exchange
fused into the kernel), "extract_batch" (loam_extract_batch: B sequences per launch), "multi_segment", "parity".

`--impl reference` times the reference's CPU implementation of the same path on the same workload (bounded sample).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "LiDAR sweeps/s scan-to-map registration"
UNIT = "sweeps/s"

# algorithmic bytes per unit of work for each kernel class (DESIGN.md "Kernels" table)
ALGO_BYTES = {
    "extract": ("28 N + 16 F per sweep", None),       # computed from counts
    "odom_knn": ("28 Q + 16 T per refresh", None),    # computed from counts
    "odom_iter": ("64 Q + 108 B per pass", 64.0),
    "to_end": ("32 B per point", 32.0),
    "map_stack": ("32 B per point", 32.0),
    "voxel": ("16 (M + V) per call", None),
    "grid": ("36 T per build", 36.0),
    "map_knn": ("96 Q per map iteration (kNN-5 + fit pair)", 96.0),
    "map_fit": ("(counted with map_knn)", 0.0),
    "gather": ("32 B per point", 32.0),
    "insert": ("32 B per point", 32.0),
    "sr_select": ("16 F + 5 N per sweep (keys, flags, labels)", 5.0 + 16.0 * 0.2),
}

# dram__bytes_read.sum + dram__bytes_write.sum per launch group of each class from the round-2 `ncu --set full` capture
# (profiles/r2_ncu_full.csv; ncu flushes caches between replays, so these are cold-cache upper bounds); None where the
# class is a chain of several kernels without one dominant launch
NCU_TRAFFIC = {"odom_knn": None,              # odom_refresh_pruned_kernel: see profiles/r2_ncu_full.csv of the final build
               "odom_iter": 386304,           # odom_loop_kernel (one launch = up to five iterations)
               "sr_select": 185344, "map_knn": 1234944, "map_fit": None, "voxel": None, "extract": None}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for k, name in enumerate(names):
                if f[3 + k].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def make_sequence(n_sweeps, rank, pinned=True):
    """All sweeps of the rank's sequence in one (pinned) host buffer + offsets in points."""
    import torch
    from gpscalibration_b200 import SweepGenerator
    gen = SweepGenerator(sensor=0, scene=0, seed=0xC0FFEE + 1000 * rank, t_offset=37.0 * rank)
    cap = gen.max_points
    host = torch.empty((n_sweeps * cap, 3), dtype=torch.float32, pin_memory=pinned)
    arr = host.numpy()
    offs = np.zeros(n_sweeps + 1, np.int64)
    for k in range(n_sweeps):
        xyz, _ = gen.sweep(k, out=arr[offs[k]:offs[k] + cap])
        offs[k + 1] = offs[k] + xyz.shape[0]
    return host, arr, offs


def reference_available():
    return all(os.path.exists(os.path.join(ROOT, "oracle", "_ref", f)) for f in ("libref_sr.so", "libref_lo.so", "libref_lm.so"))


def run_cpu(arr, offs, n_sweeps, threads3=True):
    """CPU reference path over the first n_sweeps sweeps.  Returns (seconds, kind, cores, results)."""
    if reference_available():
        from oracle import ref as refmod
        t0 = time.perf_counter()
        res = refmod.run_sequence(arr, offs[:n_sweeps + 1])
        return time.perf_counter() - t0, "reference", 3, res
    from oracle import orc
    pipe = orc.Pipeline(keep_clouds=False)
    t0 = time.perf_counter()
    if threads3:
        res = pipe.run_threaded(arr[:offs[n_sweeps]], offs[:n_sweeps + 1])
        cores = 3
    else:
        res = [pipe.process(arr[offs[k]:offs[k + 1]]) for k in range(n_sweeps)]
        cores = 1
    return time.perf_counter() - t0, "port", cores, res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-worker"])
    ap.add_argument("--worker-rank", type=int, default=0, help="(reference-worker) which rank's sequence to run")
    ap.add_argument("--sweeps", type=int, default=1000, help="sweeps per sequence (configs[1]: 1000)")
    ap.add_argument("--cpu-sweeps", type=int, default=150, help="bounded CPU sample (first sweeps of the same sequence)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--multi-segments", type=int, default=4, help="extra: independent segments sharing one GPU (0/1 = skip)")
    ap.add_argument("--no-cfg3", action="store_true", help="skip the cfg 3 object (mapping stage against a 2 M-point local map)")
    ap.add_argument("--no-cfg5", action="store_true", help="skip the cfg 5 object (1 M-point stack against the 50 M-point sharded map)")
    ap.add_argument("--no-extract-batch", action="store_true", help="skip the batched-extraction object (loam_extract_batch, B = 1 / 8 / 32)")
    ap.add_argument("--cfg5-map-points", type=int, default=50_000_000)
    ap.add_argument("--cfg5-queries", type=int, default=1_000_000)
    args = ap.parse_args()

    # exactly ONE line on stdout (the JSON): libraries that print banners to fd 1 (NCCL version line) go to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    config = {"workload": "cfg2: synthetic 1000-sweep VLP-16-shaped sequence (16x1800, reference ring table), extract + "
                          "odometry + mapping, one independent sequence per GPU",
              "sweeps_per_step": args.sweeps, "points_per_sweep": 28800, "seed": "0xC0FFEE + 1000*rank",
              "l2": "inputs larger than L2 (%d MB of sweeps per step, each touched once)" % (args.sweeps * 28800 * 12 // 1000000)}

    # ------------------------------------------------------------------------------------------- reference arm
    if args.impl == "reference-worker":  # one reference pipeline (three stage threads) on sequence `--worker-rank`
        n = min(args.cpu_sweeps, args.sweeps)
        _, arr, offs = make_sequence(n, args.worker_rank, pinned=False)
        times = []
        kind = None
        for s in range(args.warmup + args.steps):
            dt, kind, _, _ = run_cpu(arr, offs, n)
            if s >= args.warmup:
                times.append(dt)
            log(f"[reference worker {args.worker_rank}] step {s}: {n / dt:.2f} sweeps/s")
        emit({"times": times, "kind": kind})
        return 0
    if args.impl == "reference":
        if rank != 0:
            return 0
        n = min(args.cpu_sweeps, args.sweeps)
        # --gpus N: N reference pipelines side by side (N x 3 stage threads, distinct sequences: the ones the N GPU ranks
        # get), SURVEY 8d(ii) -- the reference's nodes keep their state in file-scope globals, so one process per pipeline
        env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT")}
        procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--impl", "reference-worker", "--worker-rank", str(i),
                                   "--steps", str(args.steps), "--warmup", str(args.warmup), "--sweeps", str(args.sweeps),
                                   "--cpu-sweeps", str(args.cpu_sweeps)], stdout=subprocess.PIPE, env=env, text=True)
                 for i in range(max(1, args.gpus))]
        outs = [json.loads(pr.communicate()[0].strip().splitlines()[-1]) for pr in procs]
        for pr in procs:
            if pr.returncode:
                raise RuntimeError("reference worker failed")
        kind = outs[0]["kind"]
        npipe = len(outs)
        total = max(sum(o["times"]) for o in outs)  # the slowest pipeline bounds the job, like max-over-ranks on the GPU arm
        val = npipe * n * args.steps / total
        cores = 3 * npipe
        sample = (f"first {n} sweeps per step of each of {npipe} independent sequence(s) (the ranks' own), one reference pipeline per "
                  f"sequence running concurrently, three stage threads (SR | LO | LM) each like the reference's three ROS processes; "
                  f"{os.cpu_count()} host cores present")
        emit(({"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
                          "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0

    # ------------------------------------------------------------------------------------------- our arm
    import torch
    import torch.distributed as dist
    from gpscalibration_b200 import LoamGpu

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    t_gen = time.perf_counter()
    host, arr, offs = make_sequence(args.sweeps, rank)
    n_pts = int(offs[-1])
    dev = host[:n_pts].cuda()
    log(f"[rank {rank}] generated {args.sweeps} sweeps ({n_pts * 12 / 1e6:.0f} MB) in {time.perf_counter() - t_gen:.1f}s")

    from gpscalibration_b200 import LoamGpuPipeline
    dev_t = torch.device("cuda", local_rank)
    # want_registered / want_surround: the timed arm produces /velodyne_cloud_registered (LM:1103-1112) and
    # /laser_cloud_surround (LM:1081-1101) like the reference arm always does
    gpu = LoamGpu(device=local_rank, want_registered=True, want_surround=True)          # synchronous per-sweep call (loam_process_sweep)
    pipe = LoamGpuPipeline(device=local_rank, want_registered=True, want_surround=True)  # pipelined mode: the reference's SR | LO | LM layout
    sync_stream = torch.cuda.ExternalStream(gpu.stream, device=dev_t)
    stage_streams = [torch.cuda.ExternalStream(pipe.stream(i), device=dev_t) for i in range(3)]
    join_stream = torch.cuda.Stream(device=dev_t)
    base_dev = dev.data_ptr()
    S = args.sweeps
    DEPTH = 6  # sweeps in flight before the caller starts collecting results

    def step_pipe(host_buffers):
        pipe.reset()
        last = None
        for k in range(S):
            if host_buffers:
                pipe.submit(arr[offs[k]:offs[k + 1]])
            else:
                pipe.submit_device(base_dev + int(offs[k]) * 12, int(offs[k + 1] - offs[k]))
            if k >= DEPTH:
                last = pipe.wait()
        while pipe.pending:
            last = pipe.wait()
        return last

    def step_sync(host_buffers):
        gpu.reset()
        last = None
        for k in range(S):
            if host_buffers:
                last = gpu.process_sweep(arr[offs[k]:offs[k + 1]])
            else:
                last = gpu.process_sweep_device(base_dev + int(offs[k]) * 12, int(offs[k + 1] - offs[k]))
        return last

    def timed(fn, steps, streams, stats_fn):
        """K steps bracketed by barrier + synchronize; device time from CUDA events that bracket every stage stream."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0 = stats_fn()
        w0 = time.perf_counter()
        e0.record(join_stream)
        for st in streams:
            st.wait_event(e0)
        for _ in range(steps):
            last = fn()
        for st in streams:
            ev = torch.cuda.Event()
            ev.record(st)
            join_stream.wait_event(ev)
        e1.record(join_stream)
        torch.cuda.synchronize()
        wall = time.perf_counter() - w0
        barrier()
        s1 = stats_fn()
        return e0.elapsed_time(e1), wall, {k: s1[k] - s0[k] for k in s0}, last

    for w in range(args.warmup):
        step_pipe(False)
    step_pipe(True)
    step_sync(True)
    step_sync(False)
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_dev, wall_dev, st_dev, last = timed(lambda: step_pipe(False), args.steps, stage_streams, pipe.stats)
    ms_e2e, wall_e2e, st_e2e, _ = timed(lambda: step_pipe(True), args.steps, stage_streams, pipe.stats)
    clocks = sampler.stop()
    ms_sync_dev, _, _, last_sync = timed(lambda: step_sync(False), args.steps, [sync_stream], gpu.stats)
    ms_sync_e2e, _, st_sync_e2e, _ = timed(lambda: step_sync(True), args.steps, [sync_stream], gpu.stats)
    ms_dev_max = max_over_ranks(ms_dev)
    ms_e2e_max = max_over_ranks(ms_e2e)
    value = world * S * args.steps / (ms_dev_max / 1e3)
    e2e_value = world * S * args.steps / (ms_e2e_max / 1e3)
    sync_value = world * S * args.steps / (max_over_ranks(ms_sync_dev) / 1e3)
    sync_e2e_value = world * S * args.steps / (max_over_ranks(ms_sync_e2e) / 1e3)

    # roofline: separate CUDA-event pass over one step (events around every launch group perturb the step time)
    gpu.profile(True)
    step_sync(False)
    prof = gpu.profile_read()
    gpu.profile(False)
    tot_ms = sum(v["ms"] for v in prof.values()) or 1.0
    dom = max(prof, key=lambda k: prof[k]["ms"])
    if dom == "map_fit":
        dom = "map_knn"
    d = prof[dom]
    per_unit = ALGO_BYTES[dom][1]
    if dom == "map_knn":
        d = {"ms": prof["map_knn"]["ms"] + prof["map_fit"]["ms"], "units": prof["map_knn"]["units"], "scopes": prof["map_knn"]["scopes"]}
    if per_unit is None:  # classes whose bytes depend on two sizes: use the per-sweep typical figures
        per_unit = {"extract": 28.0 + 16.0 * 0.69, "odom_knn": 28.0 + 16.0 * 2.7, "voxel": 32.0}[dom]
    algo_bytes = per_unit * d["units"]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = algo_bytes / (d["ms"] * 1e-3) / 1e9 if d["ms"] > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": NCU_TRAFFIC.get(dom), "peak_source": "MEASURED_PEAKS.json (measured)" if peaks else "fallback 6.65 TB/s",
                "algorithmic_bytes_per_unit": per_unit, "units_per_launch": d["units"] / max(1, d["scopes"]),
                "avg_launch_us": 1e3 * d["ms"] / max(1, d["scopes"]), "share_of_gpu_time": d["ms"] / tot_ms,
                "note": "single-sequence configs are launch/latency bound by construction (SURVEY §8d): ~9 MB per registration",
                "classes_ms_per_step": {k: round(v["ms"], 3) for k, v in prof.items()},
                "classes_launch_groups": {k: v["scopes"] for k, v in prof.items()}}

    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms_dev_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic", "config": config, "clocks": clocks,
           "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": st_e2e["h2d_bytes"] // args.steps,
                   "d2h_bytes_per_step": st_e2e["d2h_bytes"] // args.steps, "ms_per_step": ms_e2e_max / args.steps},
           "gpu_launches": int(st_dev["launches"]), "host_handovers_per_step": st_dev["syncs"] // args.steps,
           "wall_vs_event_ms": [round(1e3 * wall_dev, 1), round(ms_dev, 1)],
           "mode": "pipelined C ABI (loam_pipeline_submit / loam_pipeline_wait): stage threads SR | LO | LM like the reference's three processes",
           "sync_api": {"call": "loam_process_sweep (one blocking call per sweep)", "value": sync_value, "e2e": sync_e2e_value, "unit": UNIT},
           "roofline": roofline,
           "final_pose_odom": [round(float(x), 4) for x in last.odom.transform_sum],
           "results_identical_sync_vs_pipelined": list(last.odom.transform_sum) == list(last_sync.odom.transform_sum)}

    # extra (not the headline): capacity of one GPU when several independent segments share it (cfg 4 with more segments
    # than GPUs): SEG pipelined handles driven by SEG host threads over the first sweeps of the rank's sequence family
    if world == 1 and args.multi_segments > 1:
        import threading
        SEG, NS = args.multi_segments, min(400, S)
        seg_data = [(arr, offs)] + [make_sequence(NS, 100 + i, pinned=True)[1:] for i in range(1, SEG)]
        seg_pipes = [pipe] + [LoamGpuPipeline(device=local_rank, want_registered=True, want_surround=True) for _ in range(1, SEG)]
        seg_t = [0.0] * SEG

        def seg_run(i):
            a, o = seg_data[i]
            p = seg_pipes[i]
            for rep in range(2):
                p.reset()
                t0 = time.perf_counter()
                for k in range(NS):
                    p.submit(a[o[k]:o[k + 1]])
                    if k >= DEPTH:
                        p.wait()
                while p.pending:
                    p.wait()
                seg_t[i] = time.perf_counter() - t0

        ths = [threading.Thread(target=seg_run, args=(i,)) for i in range(SEG)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        out["multi_segment"] = {"segments_on_one_gpu": SEG, "sweeps_per_segment": NS, "value": SEG * NS / max(seg_t), "unit": UNIT,
                                "note": "independent sequences sharing one B200 (host buffers, wall clock of the slowest segment, second pass)"}
        for p in seg_pipes[1:]:
            p.close()
        # the same with more sequences than threads can usefully drive: EIGHT pipelines fed by ONE thread through
        # loam_pipeline_submit_batch (the extraction of the eight sweeps is one launch per kernel)
        try:
            from gpscalibration_b200 import capi as _capi
            SEGB, NSB = 8, min(300, S)
            bdata = seg_data + [make_sequence(NSB, 200 + i, pinned=True)[1:] for i in range(len(seg_data), SEGB)]
            # gn_max_ctas: the mapping loops of the eight sequences run side by side (20 CTAs each) instead of queueing for all the SMs
            bpipes = [LoamGpuPipeline(device=local_rank, want_registered=True, want_surround=True, gn_max_ctas=20) for _ in range(SEGB)]
            tb = 0.0
            for rep in range(2):
                for p in bpipes:
                    p.reset()
                t0 = time.perf_counter()
                for k in range(NSB):
                    _capi.pipeline_submit_batch(bpipes, [bdata[i][0][bdata[i][1][k]:bdata[i][1][k + 1]] for i in range(SEGB)])
                    if k >= DEPTH:
                        for p in bpipes:
                            p.wait()
                for p in bpipes:
                    while p.pending:
                        p.wait()
                tb = time.perf_counter() - t0
            out["multi_segment"]["batched_feeder"] = {"segments_on_one_gpu": SEGB, "sweeps_per_segment": NSB, "value": SEGB * NSB / tb, "unit": UNIT,
                                                      "note": "eight pipelines fed by one thread, extraction batched (loam_pipeline_submit_batch), loam_params.gn_max_ctas = 20, second pass"}
            for p in bpipes:
                p.close()
        except Exception as e:
            out["multi_segment"]["batched_feeder"] = {"error": repr(e)}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n = min(args.cpu_sweeps, S)
        dt, kind, cores, cpu_res = run_cpu(arr, offs, n)
        out["cpu_baseline"] = {"value": n / dt, "unit": UNIT, "cores": cores, "kind": kind,
                               "sample": f"first {n} sweeps of the same sequence, three stage threads (SR | LO | LM), {os.cpu_count()} host cores present"}
        # trajectory parity (SURVEY 8d cfg 2): the CUDA path's poses over the same sweeps against the CPU arm's, bit for bit
        # (a fresh handle: the publish cadence LO:1099-1106 depends on frameCount, which survives a reset -- the CPU arm starts fresh too)
        g2 = LoamGpu(device=local_rank, want_registered=True, want_surround=True)
        worst_o = worst_m = 0.0
        n_map = 0
        for k in range(n):
            r = g2.process_sweep(arr[offs[k]:offs[k + 1]])
            if kind == "reference":
                co, cm = cpu_res[k]
            else:
                co, cm = np.array(cpu_res[k].odom), (np.array(cpu_res[k].mapped) if cpu_res[k].mapping_ran else None)
            worst_o = max(worst_o, float(np.abs(np.array(r.odom.transform_sum, np.float32) - np.asarray(co, np.float32)).max()))
            assert bool(r.mapping_ran) == (cm is not None), k
            if cm is not None:
                n_map += 1
                worst_m = max(worst_m, float(np.abs(np.array(r.map.transform_aft_mapped, np.float32) - np.asarray(cm, np.float32)).max()))
        g2.close()
        out["parity"] = {"sweeps": n, "mapping_runs": n_map, "against": kind, "max_abs_pose_diff_odometry": worst_o,
                         "max_abs_pose_diff_mapping": worst_m}
    # cfg 1 (configs[0]): ONE registration of a sweep pair (extract + scan-to-scan odometry), the case that can only be
    # latency bound: microseconds per stage against the floor the launches and host hand-overs of that stage set
    if world == 1:
        g1 = LoamGpu(device=local_rank)
        st1 = torch.cuda.ExternalStream(g1.stream, device=dev_t)
        launch_period_us, roundtrip_us = g1.launch_latency(2000)  # empty kernel: back-to-back period, launch + host-visible completion
        reg_us, ext_us, iters, launches_reg, launches_ext, syncs_reg = [], [], 0, 0, 0, 0
        for rep in range(12):
            g1.reset()
            for k in range(2):  # sweep 0 initialises (LO:519-563); sweep 1 finds laserCloudCornerLastNum == 0 and only
                g1.extract(arr[offs[k]:offs[k + 1]])  # hands its clouds on (LO:572, 1119-1121); sweep 2 is the first registration
                g1.odometry_process()
            s_a = g1.stats()
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record(st1)
            g1.extract(arr[offs[2]:offs[3]])
            s_b = g1.stats()
            e1.record(st1)
            ro = g1.odometry_process()
            e2.record(st1)
            torch.cuda.synchronize()
            s_c = g1.stats()
            if rep >= 2:
                ext_us.append(1e3 * e0.elapsed_time(e1))
                reg_us.append(1e3 * e1.elapsed_time(e2))
            iters = ro.iterations
            launches_ext, launches_reg = s_b["launches"] - s_a["launches"], s_c["launches"] - s_b["launches"]
            syncs_reg = s_c["syncs"] - s_b["syncs"]
        g1.close()
        refreshes = (iters + 4) // 5  # LO:595: correspondences are re-searched every fifth iteration
        floor_us = launches_reg * launch_period_us + syncs_reg * roundtrip_us
        out["cfg1"] = {"workload": "cfg1: one VLP-16-shaped sweep pair (28.8 k points each): extract of the second sweep + its scan-to-scan registration against the first",
                       "extract_us": float(np.median(ext_us)), "extract_launches": int(launches_ext),
                       "registration_us": float(np.median(reg_us)), "iterations": int(iters), "refreshes": int(refreshes),
                       "registration_launches": int(launches_reg), "registration_host_handovers": int(syncs_reg),
                       "us_per_iteration": float(np.median(reg_us)) / max(1, iters),
                       "launch_period_us": launch_period_us, "host_roundtrip_us": roundtrip_us,
                       "latency_floor_us": floor_us, "registration_over_floor": float(np.median(reg_us)) / floor_us if floor_us > 0 else None,
                       "note": "floor = launches x back-to-back launch period + host hand-overs x (launch + host-visible completion), both measured here"}
    gpu.close()
    pipe.close()
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    if world == 1 and not args.no_cfg3:
        import bench_cfg3
        try:
            out["cfg3"] = bench_cfg3.run_stage(local=local_rank, log=log)
        except Exception as e:  # an extra must not take the headline down
            out["cfg3"] = {"error": repr(e)}
    if world == 1 and not args.no_extract_batch:
        import bench_extract_batch
        try:  # SURVEY 8b `*_batch`: the extraction stage with one launch per kernel for B sequences (roofline on batched launches)
            out["extract_batch"] = bench_extract_batch.run((1, 8, 32), 24, local_rank, peak_gbs, log=log)
        except Exception as e:
            out["extract_batch"] = {"error": repr(e)}
    if not args.no_cfg5:
        import bench_cfg5
        try:
            c5 = bench_cfg5.run(args.cfg5_map_points, args.cfg5_queries, 10, rank, local_rank, world, dist if world > 1 else None, log=log)
            if "fused" in c5:
                c5["fused"]["frac_of_hbm_peak_per_gpu"] = c5["fused"]["per_gpu_GBps_algorithmic"] / peak_gbs
            out["cfg5"] = c5
        except Exception as e:
            out["cfg5"] = {"error": repr(e)}
    if rank == 0:
        emit(out)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
