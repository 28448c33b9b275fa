"""Throughput of the pipelined mode vs the fused synchronous call (diagnostic, run under gpurun)."""
import sys, time
sys.path.insert(0, '.')
from gpscalibration_b200 import LoamGpu, LoamGpuPipeline, SweepGenerator
N = int(sys.argv[1]) if len(sys.argv) > 1 else 300
gen = SweepGenerator()
sw = [gen.sweep(k)[0].copy() for k in range(N)]
g = LoamGpu(want_registered=(len(sys.argv) > 2 and sys.argv[2] == 'full'), want_surround=(len(sys.argv) > 2 and sys.argv[2] == 'full'))
for rep in range(2):
    g.reset(); t0 = time.time(); ref = [g.process_sweep(x) for x in sw]; t = time.time() - t0
print("fused     %.3f s  %.1f sweeps/s" % (t, N / t))
WANT = len(sys.argv) > 2 and sys.argv[2] == 'full'
p = LoamGpuPipeline(want_registered=WANT, want_surround=WANT)
for rep in range(3):
    p.reset()
    p.stage_times()
    for w in range(3):
        p.stage_host_times(w)
    t0 = time.time()
    got = []
    for k, x in enumerate(sw):
        p.submit(x)
        if k >= 6:
            got.append(p.wait())
    while p.pending:
        got.append(p.wait())
    t = time.time() - t0
    print("pipelined %.3f s  %.1f sweeps/s" % (t, N / t), "stage busy ms/sweep", [round(1e3 * x / N, 4) for x in p.stage_times()])
    for w in range(3):
        print("   stage", w, "host sections us/sweep", [round(1e6 * x / N, 1) for x in p.stage_host_times(w)])
same = all(list(a.odom.transform_sum) == list(b.odom.transform_sum) for a, b in zip(ref, got))
print("identical odometry:", same, "final mapped", list(got[-1].map.transform_aft_mapped), list(ref[-1].map.transform_aft_mapped))
p.close()
