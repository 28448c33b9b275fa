"""Long-run parity + host-time breakdown (diagnostic script run under gpurun)."""
import sys, time, json
import numpy as np
sys.path.insert(0, '.')
from gpscalibration_b200 import LoamGpu, SweepGenerator
from oracle import orc
N = int(sys.argv[1]) if len(sys.argv) > 1 else 200
gen = SweepGenerator()
sw = [gen.sweep(k)[0].copy() for k in range(N)]
gpu = LoamGpu()
pipe = orc.Pipeline(keep_clouds=False)
worst_o = worst_m = 0.0
first_bad = None
t0 = time.time()
res = []
for k in range(N):
    res.append(gpu.process_sweep(sw[k]))
tg = time.time() - t0
print("gpu %.2f s (%.1f sweeps/s)" % (tg, N / tg))
print("host times", json.dumps({k: round(v, 4) for k, v in gpu.host_times().items()}))
print("stats", gpu.stats())
t0 = time.time()
for k in range(N):
    o = pipe.process(sw[k])
    r = res[k]
    do = np.abs(np.array(r.odom.transform_sum) - np.array(o.odom)).max()
    worst_o = max(worst_o, do)
    if r.mapping_ran:
        dm = np.abs(np.array(r.map.transform_aft_mapped) - np.array(o.mapped)).max()
        worst_m = max(worst_m, dm)
        if (r.map.n_corner_map, r.map.n_surf_map, r.map.iterations) != (o.n_corner_map, o.n_surf_map, o.map_iters) and first_bad is None:
            first_bad = (k, r.map.n_corner_map, o.n_corner_map, r.map.n_surf_map, o.n_surf_map, r.map.iterations, o.map_iters)
    if (do > 0 or r.odom.iterations != o.odom_iters) and first_bad is None:
        first_bad = (k, 'odom', do, r.odom.iterations, o.odom_iters)
print("cpu %.2f s" % (time.time() - t0))
print("worst odom diff", worst_o, "worst map diff", worst_m, "first mismatch", first_bad)
print("final odom", list(res[-1].odom.transform_sum), "final mapped", list(res[-2].map.transform_aft_mapped), list(res[-1].map.transform_aft_mapped))
