"""Batched extraction (loam_extract_batch): B independent VLP-16-shaped sequences, one sweep each per call,
every extraction kernel launched once for the whole batch (grid.y = sequence).  Reports sweeps/s of the extraction stage
and the achieved fraction of the HBM roofline on its ALGORITHMIC bytes (28 N + 16 F per sweep) against the same
sweeps through B separate loam_extract calls.  Host buffers (pinned), H2D inside the timed region.
`run()` returns the dict bench.py embeds under "extract_batch"."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402


def run(batches=(1, 8, 32), sweeps_per_seq=24, local=0, peak_gbs=6553.9, log=lambda *a: None):
    import torch
    from gpscalibration_b200 import LoamGpu, SweepGenerator, capi
    Bmax = max(batches)
    gens = [SweepGenerator(sensor=0, scene=b % 2, seed=0xC0FFEE + 1000 * b, t_offset=37.0 * b) for b in range(min(Bmax, 8))]
    # eight distinct sequences, reused round-robin for larger batches (the cost does not depend on the content)
    base = [[g.sweep(k)[0] for k in range(sweeps_per_seq)] for g in gens]
    pinned = []
    for b in range(Bmax):
        row = []
        for k in range(sweeps_per_seq):
            x = base[b % len(base)][k]
            t = torch.empty(x.shape, dtype=torch.float32, pin_memory=True)
            t.numpy()[:] = x
            row.append(t.numpy())
        pinned.append(row)
    handles = [LoamGpu(device=local) for _ in range(Bmax)]
    out = {"workload": "extraction stage only: one VLP-16-shaped sweep (28.8 k points) per sequence and call, host buffers",
           "algorithmic_bytes": "28 N + 16 F per sweep", "hbm_peak_gbs": peak_gbs, "points": []}
    for B in batches:
        hs = handles[:B]
        res = {}
        for mode in ("single", "batch"):
            best = None
            feats = 0
            for rep in range(3):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for k in range(sweeps_per_seq):
                    if mode == "batch":
                        cs = capi.extract_batch(hs, [pinned[b][k] for b in range(B)])
                    else:
                        cs = [hs[b].extract(pinned[b][k]) for b in range(B)]
                    if rep == 0:
                        feats += sum(c.n_sharp + c.n_less_sharp + c.n_flat + c.n_less_flat for c in cs)
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
                if rep == 0:
                    res[mode + "_features"] = feats
            res[mode] = best
        n_sw = B * sweeps_per_seq
        n_pts = sum(pinned[b][k].shape[0] for b in range(B) for k in range(sweeps_per_seq))
        byts = 28.0 * n_pts + 16.0 * res["batch_features"]
        pt = {"sequences": B, "single_calls_sweeps_per_s": n_sw / res["single"], "batched_sweeps_per_s": n_sw / res["batch"],
              "batched_us_per_call": 1e6 * res["batch"] / sweeps_per_seq, "speedup": res["single"] / res["batch"],
              "batched_algorithmic_GBps": byts / res["batch"] / 1e9, "frac_of_hbm_peak": byts / res["batch"] / 1e9 / peak_gbs,
              "launches_per_sweep_batched": 8.0 / B}
        log(f"[extract_batch] B={B}: single {pt['single_calls_sweeps_per_s']:.0f} sweeps/s, batched {pt['batched_sweeps_per_s']:.0f} "
            f"({pt['batched_us_per_call']:.0f} us per call, {pt['batched_algorithmic_GBps']:.1f} GB/s)")
        out["points"].append(pt)
    for h in handles:
        h.close()
    return out


if __name__ == "__main__":
    bs = tuple(int(a) for a in sys.argv[1:]) or (1, 2, 4, 8, 16, 32, 64)
    print(json.dumps(run(bs, log=lambda *a: print(*a, file=sys.stderr, flush=True))))
