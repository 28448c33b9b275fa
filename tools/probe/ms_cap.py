"""8 pipelines fed with batched extraction, for the LOAM_GN_MAX_CTAS experiment: python tools/probe/ms_cap.py [S]"""
import sys, time
sys.path.insert(0, '.')
from gpscalibration_b200 import LoamGpuPipeline, SweepGenerator, capi
S = int(sys.argv[1]) if len(sys.argv) > 1 else 8
N = 300
seqs = []
for s in range(S):
    gen = SweepGenerator(seed=0xC0FFEE + 1000 * s, t_offset=37.0 * s)
    seqs.append([gen.sweep(k)[0].copy() for k in range(N)])
objs = [LoamGpuPipeline(want_registered=True, want_surround=True) for _ in range(S)]
for rep in range(2):
    for p in objs: p.reset()
    t0 = time.perf_counter()
    for k in range(N):
        capi.pipeline_submit_batch(objs, [seqs[i][k] for i in range(S)])
        if k >= 6:
            for p in objs: p.wait()
    for p in objs:
        while p.pending: p.wait()
    dt = time.perf_counter() - t0
print("S", S, "aggregate %.0f sweeps/s" % (S * N / dt))
