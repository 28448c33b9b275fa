"""Phase stamps of the mapping Gauss-Newton kernel inside the real pipeline (cfg 2 sizes): LOAM_GN_DEBUG=1 python tools/probe/gn_small.py"""
import sys
sys.path.insert(0, '.')
from gpscalibration_b200 import LoamGpu, SweepGenerator
gen = SweepGenerator()
g = LoamGpu()
for k in range(60):
    g.process_sweep(gen.sweep(k)[0].copy())
g.close()
