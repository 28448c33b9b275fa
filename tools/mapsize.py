import sys
sys.path.insert(0, '.')
from gpscalibration_b200 import LoamGpu, SweepGenerator
gen = SweepGenerator()
g = LoamGpu()
for k in range(300):
    r = g.process_sweep(gen.sweep(k)[0].copy())
    if k % 50 == 49 or k < 4:
        m = r.map
        print(k, m.n_corner_stack, m.n_surf_stack, m.n_corner_map, m.n_surf_map, m.iterations, r.odom.iterations)
