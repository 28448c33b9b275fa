"""Aggregate throughput of S independent sequences sharing ONE GPU (cfg 4 with more segments than GPUs):
S pipelined handles driven by S host threads (ctypes releases the GIL)."""
import sys, time, threading
sys.path.insert(0, '.')
import numpy as np
from gpscalibration_b200 import LoamGpuPipeline, SweepGenerator
N = int(sys.argv[1]) if len(sys.argv) > 1 else 400
for S in (1, 2, 4, 8):
    seqs = []
    for s in range(S):
        gen = SweepGenerator(seed=0xC0FFEE + 1000 * s, t_offset=37.0 * s)
        seqs.append([gen.sweep(k)[0].copy() for k in range(N)])
    pipes = [LoamGpuPipeline() for _ in range(S)]
    def run(p, sw, out):
        for rep in range(2):
            p.reset()
            t0 = time.perf_counter()
            for k, x in enumerate(sw):
                p.submit(x)
                if k >= 6:
                    p.wait()
            while p.pending:
                p.wait()
            out.append(time.perf_counter() - t0)
    outs = [[] for _ in range(S)]
    ths = [threading.Thread(target=run, args=(pipes[i], seqs[i], outs[i])) for i in range(S)]
    t0 = time.perf_counter()
    for t in ths: t.start()
    for t in ths: t.join()
    wall = time.perf_counter() - t0
    last = max(o[1] for o in outs)
    print("segments %d: second pass %.3f s -> aggregate %.0f sweeps/s (per segment %.0f)" % (S, last, S * N / last, N / last), flush=True)
    for p in pipes: p.close()
