"""Parity sweep over generator settings the unit tests do not use (diagnostic, run under gpurun): for every (sensor,
scene, seed) the CUDA pipeline and the CPU oracle pipeline must agree bit for bit on poses and iteration counts."""
import sys
import numpy as np
sys.path.insert(0, '.')
from gpscalibration_b200 import LoamGpu, SweepGenerator
from oracle import orc

N = int(sys.argv[1]) if len(sys.argv) > 1 else 24
bad = 0
for sensor, ring in ((0, dict(n_scans=16)), (2, dict(n_scans=64, ring_mode=1, ring_ang_min=-24.8, ring_ang_step=26.8 / 63.0))):
    for scene in (0, 1):
        for seed in (1, 77, 4242):
            gen = SweepGenerator(sensor=sensor, scene=scene, seed=seed)
            gpu = LoamGpu(**ring)
            pipe = orc.Pipeline(ring.get("n_scans", 16), ring.get("ring_mode", 0), ring.get("ring_ang_min", -15.0), ring.get("ring_ang_step", 2.0))
            n = N if sensor == 0 else max(6, N // 4)
            ok = True
            for k in range(n):
                xyz = gen.sweep(k)[0]
                r, o = gpu.process_sweep(xyz), pipe.process(xyz)
                same = np.array_equal(np.array(r.odom.transform_sum, np.float32), np.array(o.odom, np.float32)) and \
                    (not o.odom_published or r.odom.iterations == o.odom_iters) and r.mapping_ran == o.mapping_ran
                if r.mapping_ran:
                    same = same and np.array_equal(np.array(r.map.transform_aft_mapped, np.float32), np.array(o.mapped, np.float32)) and \
                        r.map.iterations == o.map_iters and (r.map.n_corner_map, r.map.n_surf_map) == (o.n_corner_map, o.n_surf_map)
                if not same:
                    ok = False
                    print("MISMATCH", sensor, scene, seed, k, list(r.odom.transform_sum), list(o.odom), r.odom.iterations, o.odom_iters)
                    break
            bad += 0 if ok else 1
            print("sensor", sensor, "scene", scene, "seed", seed, "sweeps", n, "OK" if ok else "FAILED")
            gpu.close()
print("failures:", bad)
sys.exit(1 if bad else 0)
