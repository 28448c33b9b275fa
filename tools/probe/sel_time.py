"""Per-class device time of the extraction stage over N sweeps (loam_profile CUDA-event timers): sr_select vs the rest."""
import sys, time
sys.path.insert(0, '.')
import os
from gpscalibration_b200 import LoamGpu, SweepGenerator, capi
if os.environ.get("LOAM_LIB"):  # a debug build of the library (-DLG_SEL_DEBUG: phase stamps of sr_select_kernel)
    capi.library_path = lambda: os.path.abspath(os.environ["LOAM_LIB"])
N = int(sys.argv[1]) if len(sys.argv) > 1 else 200
gen = SweepGenerator()
sw = [gen.sweep(k)[0].copy() for k in range(N)]
gpu = LoamGpu()
for k in range(20):
    gpu.extract(sw[k])
gpu.profile(True)
t0 = time.perf_counter()
for k in range(N):
    gpu.extract(sw[k])
dt = time.perf_counter() - t0
p = gpu.profile_read()
print(f"extract wall {dt / N * 1e6:.1f} us per sweep (profiling on)")
for k, v in p.items():
    if v["scopes"]:
        print(f"  {k:10s} {v['ms'] / v['scopes'] * 1e3:8.2f} us per scope, {v['scopes']} scopes")
try:
    import ctypes as C
    out = (C.c_longlong * 8)()
    gpu.lib.loam_debug_sel(out, 0)
    k = max(out[7], 1)
    names = ["setup", "ranks", "walk-init", "walk-rounds", "walk-number", "dominators+tail"]
    print("ring-1 CTA cycles per launch: " + ", ".join(f"{n} {out[i] / k:.0f}" for i, n in enumerate(names)) + f"; rounds {out[6] / k:.1f}")
except AttributeError:
    pass
