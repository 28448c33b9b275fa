import sys, time
sys.path.insert(0, '.')
from gpscalibration_b200 import LoamGpuPipeline, SweepGenerator, capi
S, N = 12, 300
seqs = []
for s in range(S):
    gen = SweepGenerator(seed=0xC0FFEE + 1000 * s, t_offset=37.0 * s)
    seqs.append([gen.sweep(k)[0].copy() for k in range(N)])
objs = [LoamGpuPipeline(want_registered=True, want_surround=True) for _ in range(S)]
try:
  for rep in range(2):
    for p in objs: p.reset()
    for k in range(N):
        capi.pipeline_submit_batch(objs, [seqs[i][k] for i in range(S)], lockstep=True)
        if k >= 6:
            for p in objs: p.wait()
    for p in objs:
        while p.pending: p.wait()
    print("rep ok", rep)
except Exception as e:
    print("ERR", e, capi.load_library().loam_last_cuda_error(None))
