"""Steady-state host-time / kernel-class breakdown of the fused pipeline (diagnostic, run under gpurun)."""
import sys, time, json
import numpy as np
sys.path.insert(0, '.')
from gpscalibration_b200 import LoamGpu, SweepGenerator
N = int(sys.argv[1]) if len(sys.argv) > 1 else 300
gen = SweepGenerator()
sw = [gen.sweep(k)[0].copy() for k in range(N)]
WANT = (len(sys.argv) > 2 and sys.argv[2] == "full")
gpu = LoamGpu(want_registered=WANT, want_surround=WANT)
for rep in range(3):
    gpu.reset()
    gpu.host_times()
    s0 = gpu.stats()
    if rep == 2:
        gpu.profile(True)
    t0 = time.time()
    its = mits = 0
    for k in range(N):
        r = gpu.process_sweep(sw[k])
        its += r.odom.iterations
        mits += r.map.iterations if r.mapping_ran else 0
    tg = time.time() - t0
    s1 = gpu.stats()
    print("rep %d: %.3f s (%.1f sweeps/s) odom iters %d map iters %d" % (rep, tg, N / tg, its, mits))
    print("  host ms/sweep", json.dumps({k: round(1e3 * v / N, 3) for k, v in gpu.host_times().items()}))
    print("  per sweep", {k: (s1[k] - s0[k]) / N for k in s0})
    if rep == 2:
        p = gpu.profile_read()
        print("  gpu ms/sweep", json.dumps({k: round(v["ms"] / N, 4) for k, v in p.items()}))
        print("  gpu us/scope", json.dumps({k: round(1e3 * v["ms"] / max(1, v["scopes"]), 1) for k, v in p.items()}))
