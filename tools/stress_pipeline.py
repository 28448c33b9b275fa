"""Pipelined mode against the blocking call over long sequences (diagnostic, run under gpurun): every field of every sweep's
result (poses, iteration counts, cloud sizes, registered / surround counts) must be equal, across resets in mid-stream and
with several pipelines sharing the GPU -- the output stage, the deferred cube read-back and the second mapping stream
must never change a result.  usage: stress_pipeline.py [sweeps] [pipelines]"""
import sys, threading
import numpy as np
sys.path.insert(0, '.')
from gpscalibration_b200 import LoamGpu, LoamGpuPipeline, SweepGenerator

N = int(sys.argv[1]) if len(sys.argv) > 1 else 600
P = int(sys.argv[2]) if len(sys.argv) > 2 else 3
RESETS = {N // 3, N // 3 + 1, (2 * N) // 3}  # reset before these sweeps (two in a row: a one-sweep epoch)


def key(r):
    o, m = r.odom, r.map
    t = (r.counts.n_full, r.counts.n_sharp, r.counts.n_less_sharp, r.counts.n_flat, r.counts.n_less_flat, list(o.transform_sum), o.iterations,
         o.odom_published, o.clouds_published, o.fullres_published, r.mapping_ran)
    if r.mapping_ran:
        t += (list(m.transform_aft_mapped), list(m.transform_bef_mapped), m.iterations, m.n_corner_stack, m.n_surf_stack, m.n_corner_map,
              m.n_surf_map, m.surround_published, m.n_surround, m.n_registered)
    return t


def one(idx, out):
    gen = SweepGenerator(sensor=0, scene=idx % 2, seed=0xC0FFEE + 1000 * idx, t_offset=37.0 * idx)
    sw = [gen.sweep(k)[0].copy() for k in range(N)]
    a = LoamGpu(want_registered=True, want_surround=True)
    ref = []
    for k, x in enumerate(sw):
        if k in RESETS:
            a.reset()
        ref.append(key(a.process_sweep(x)))
    a.close()
    p = LoamGpuPipeline(want_registered=True, want_surround=True)
    got, pending = [], 0
    for k, x in enumerate(sw):
        if k in RESETS:
            p.reset()
        p.submit(x)
        pending += 1
        if pending > 6:
            got.append(key(p.wait()))
            pending -= 1
    while p.pending:
        got.append(key(p.wait()))
    p.close()
    bad = [k for k in range(N) if ref[k] != got[k]]
    n_sur = sum(1 for r in ref if r[10] and r[18])
    out[idx] = (bad[:3], n_sur, ref[bad[0]] if bad else None, got[bad[0]] if bad else None)


outs = {}
ths = [threading.Thread(target=one, args=(i, outs)) for i in range(P)]
for t in ths: t.start()
for t in ths: t.join()
fail = 0
for i in range(P):
    bad, n_sur, r, g = outs[i]
    print("pipeline", i, "sweeps", N, "surround runs", n_sur, "OK" if not bad else ("MISMATCH at %s\n  blocking %s\n  pipelined %s" % (bad, r, g)))
    fail += 1 if bad else 0
print("failures:", fail)
sys.exit(1 if fail else 0)
