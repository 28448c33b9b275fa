"""Aggregate throughput of S independent sequences sharing ONE GPU (cfg 4 with more segments than GPUs), three ways:
  pipelined  S pipelined handles (3 stage threads + 3 streams each) driven by S host threads
  batched    the same pipelines fed by ONE thread through loam_pipeline_submit_batch (extraction of the S sweeps batched)
  lockstep   ... through loam_pipeline_submit_lockstep (extraction AND odometry batched in lock-step, mapping per pipeline)
  blocking   S plain handles, one host thread each calling loam_process_sweep (1 thread + 1 stream per sequence)
ctypes releases the GIL during the calls.  usage: multiseg.py [sweeps] [full]"""
import sys, time, threading, os
sys.path.insert(0, '.')
import numpy as np
from gpscalibration_b200 import LoamGpu, LoamGpuPipeline, SweepGenerator
N = int(sys.argv[1]) if len(sys.argv) > 1 else 400
WANT = len(sys.argv) > 2 and sys.argv[2] == "full"
print("host cores", os.cpu_count(), "registered+surround", WANT, flush=True)
seqs_all = []
for s in range(16):
    gen = SweepGenerator(seed=0xC0FFEE + 1000 * s, t_offset=37.0 * s)
    seqs_all.append([gen.sweep(k)[0].copy() for k in range(N)])
for mode in ("pipelined", "batched", "lockstep", "blocking"):
    for S in (1, 2, 4, 8, 12, 16):
        seqs = seqs_all[:S]
        if mode in ("pipelined", "batched", "lockstep"):
            cap = 20 if (S > 1 and mode != "pipelined") else 0  # loam_params.gn_max_ctas: mapping loops side by side
            objs = [LoamGpuPipeline(want_registered=WANT, want_surround=WANT, gn_max_ctas=cap) for _ in range(S)]
        else:
            objs = [LoamGpu(want_registered=WANT, want_surround=WANT) for _ in range(S)]
        def run(p, sw, out):
            for rep in range(2):
                p.reset()
                t0 = time.perf_counter()
                if mode == "pipelined":
                    for k, x in enumerate(sw):
                        p.submit(x)
                        if k >= 6:
                            p.wait()
                    while p.pending:
                        p.wait()
                else:
                    for x in sw:
                        p.process_sweep(x)
                out.append(time.perf_counter() - t0)
        outs = [[] for _ in range(S)]
        if mode == "lockstep" and S >= 4:  # lock-step groups: G feeder threads, S / G pipelines each (rounds of different groups overlap)
            from gpscalibration_b200 import capi
            G = 2 if S < 8 else 4
            groups = [list(range(g, S, G)) for g in range(G)]
            def feed(idx, res):
                for rep in range(2):
                    for i in idx: objs[i].reset()
                    t0 = time.perf_counter()
                    for k in range(N):
                        capi.pipeline_submit_batch([objs[i] for i in idx], [seqs[i][k] for i in idx], lockstep=True)
                        if k >= 6:
                            for i in idx: objs[i].wait()
                    for i in idx:
                        while objs[i].pending: objs[i].wait()
                    res.append(time.perf_counter() - t0)
            gres = [[] for _ in range(G)]
            ths = [threading.Thread(target=feed, args=(groups[g], gres[g])) for g in range(G)]
            for t in ths: t.start()
            for t in ths: t.join()
            for o in outs: o.extend([max(r[0] for r in gres), max(r[1] for r in gres)])
        elif mode in ("batched", "lockstep"):  # ONE feeder thread: loam_pipeline_submit_batch extracts the S sweeps with one launch per kernel
            from gpscalibration_b200 import capi
            for rep in range(2):
                for p in objs: p.reset()
                t0 = time.perf_counter()
                for k in range(N):
                    capi.pipeline_submit_batch(objs, [seqs[i][k] for i in range(S)], lockstep=(mode == "lockstep"))
                    if k >= 6:
                        for p in objs: p.wait()
                for p in objs:
                    while p.pending: p.wait()
                dt = time.perf_counter() - t0
                for o in outs: o.append(dt)
        else:
            ths = [threading.Thread(target=run, args=(objs[i], seqs[i], outs[i])) for i in range(S)]
            for t in ths: t.start()
            for t in ths: t.join()
        last = max(o[1] for o in outs)
        print("%s segments %2d: second pass %.3f s -> aggregate %6.0f sweeps/s (per segment %.0f)" % (mode, S, last, S * N / last, N / last), flush=True)
        for p in objs: p.close()
