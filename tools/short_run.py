"""Short fused-pipeline run used as the ncu target (N sweeps, host buffers; registered + surround outputs on)."""
import sys
sys.path.insert(0, '.')
from gpscalibration_b200 import LoamGpu, SweepGenerator
N = int(sys.argv[1]) if len(sys.argv) > 1 else 40
gen = SweepGenerator()
sw = [gen.sweep(k)[0].copy() for k in range(N)]
gpu = LoamGpu(want_registered=True, want_surround=True)
for k in range(N):
    r = gpu.process_sweep(sw[k])
print("ok", list(r.odom.transform_sum), gpu.stats())
