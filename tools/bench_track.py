"""N4 track calibration: the LD:57-83 loop (6 weighted alignments + 6 smoothing passes) on the GPU-exact path, the host
closed form and -- where oracle/_ref/libref_tc.so travelled -- the reference's own O(N^2) code (diagnostic, run under gpurun)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np  # noqa: E402
from gpscalibration_b200 import capi  # noqa: E402
from test_track_calibration import make_tracks  # noqa: E402

out = []
for n in [int(a) for a in sys.argv[1:]] or [2000, 20000, 100000]:
    slam, enu = make_tracks(n, 5, outliers=n // 50)
    r = {"track_points": n}
    for name, mode in (("gpu_exact", 0), ("host_closed_form", 1)):
        capi.track_calibrate_long(slam, enu, 5, mode=mode)
        t0 = time.perf_counter()
        w, cal = capi.track_calibrate_long(slam, enu, 5, mode=mode)
        r[name + "_ms"] = 1e3 * (time.perf_counter() - t0)
        r[name] = cal
    r["closed_form_max_abs_diff_m"] = float(np.abs(r.pop("gpu_exact") - r.pop("host_closed_form")).max())
    r["pair_updates_per_s_gpu"] = 6.0 * n * n / (r["gpu_exact_ms"] * 1e-3)
    try:
        from oracle import ref
        if ref.tc_available() and n <= 20000:
            t0 = time.perf_counter()
            w2, cal2 = ref.tc_long(slam, enu, 5)
            r["reference_cpu_ms"] = 1e3 * (time.perf_counter() - t0)
            r["gpu_equals_reference"] = bool(np.array_equal(cal2, cal if mode == 0 else cal2) and np.array_equal(capi.track_calibrate_long(slam, enu, 5, mode=0)[1], cal2))
    except Exception as e:  # the oracle is optional here
        r["reference_cpu_ms"] = repr(e)
    out.append(r)
print(json.dumps(out))
