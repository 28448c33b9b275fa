"""Map Gauss-Newton kernel layouts side by side on a synthetic voxel-filtered map: time per iteration of the device loop
and equality of the final pose.  usage: python tools/probe/gn_modes.py [map_points] [queries]"""
import os, subprocess, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

def one(mp, nq):
    import torch
    from gpscalibration_b200 import LoamGpu, mapsynth
    T_TRUE = np.array([0.002, 0.01, -0.003, 0.05, -0.02, 0.08], np.float32) * float(os.environ.get("T_SCALE", "1"))
    cm, sm, E = mapsynth.synth_map(mp - mp // 5, mp // 5)
    cs, ss = mapsynth.synth_queries(cm, sm, nq, T_TRUE)
    gpu = LoamGpu(device=0)
    dev = torch.device("cuda", 0)
    st = torch.cuda.ExternalStream(gpu.stream, device=dev)
    gpu.map_set_inputs(cs, ss, cm, sm)
    ms = []
    for rep in range(5):
        gpu.map_set_inputs(cs, ss, cm, sm)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        Tf, it = gpu.map_optimize(np.zeros(6, np.float32), 10)
        e1.record(st)
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    part = torch.zeros(32, dtype=torch.float64, device=dev)
    gpu.map_iter_partial(0, np.zeros(6, np.float32), part.data_ptr())
    sums = part[:28].cpu().numpy()
    lc, ls = gpu.map_corr(cs.shape[0], ss.shape[0])
    import hashlib
    print(json.dumps({"sub": os.environ.get("LOAM_GN_SUB", "auto"), "t_scale": os.environ.get("T_SCALE"), "ms": [round(x, 3) for x in ms], "map": int(cm.shape[0] + sm.shape[0]),
                      "nq": int(cs.shape[0] + ss.shape[0]), "iters": int(it), "loop_ms": float(np.median(ms[1:])),
                      "iter_us": 1e3 * float(np.median(ms[1:])) / it, "T": [float(x) for x in Tf], "n_sel": int(sums[27]),
                      "sum0": float(sums[0]), "nbr_sha": hashlib.sha1(lc.tobytes() + ls.tobytes()).hexdigest()[:12]}), flush=True)
    gpu.close()

if __name__ == "__main__":
    if len(sys.argv) > 3 and sys.argv[3] == "one":
        one(int(sys.argv[1]), int(sys.argv[2]))
    else:
        mp = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
        nq = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
        for env in ({"LOAM_GN_SUB": "8"}, {"LOAM_GN_SUB": "1"}):
            e = dict(os.environ); e.update(env)
            subprocess.run([sys.executable, os.path.abspath(__file__), str(mp), str(nq), "one"], env=e)
