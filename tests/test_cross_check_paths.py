"""The two diagnostic switches keep the older, simpler code paths alive as cross-checks:
  LOAM_ODOM_BRUTE_FORCE=1  brute-force nearest neighbour + literal ring scans instead of the box-pruned kernels
  LOAM_HOST_GN_LOOP=1      every odometry Gauss-Newton iteration through the host instead of the device loop
Both are read once per process, so each variant runs in its own interpreter; all must produce the same poses bit for bit."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPT = r"""
import sys
sys.path.insert(0, %r)
import numpy as np
from gpscalibration_b200 import LoamGpu, SweepGenerator
gen = SweepGenerator()
gpu = LoamGpu()
out = []
for k in range(36):
    r = gpu.process_sweep(gen.sweep(k)[0].copy())
    out.append(np.array(list(r.odom.transform_sum) + list(r.map.transform_aft_mapped) + [r.odom.iterations, r.map.iterations], np.float32))
print(np.stack(out).tobytes().hex())
""" % ROOT


def _run(env_extra):
    env = dict(os.environ)
    env.pop("LOAM_ODOM_BRUTE_FORCE", None)
    env.pop("LOAM_HOST_GN_LOOP", None)
    env.update(env_extra)
    p = subprocess.run([sys.executable, "-c", SCRIPT], env=env, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    return p.stdout.strip().splitlines()[-1]


@pytest.mark.gpu
def test_pruned_and_device_loop_equal_their_cross_checks(_built):
    base = _run({})
    assert len(base) > 1000
    assert _run({"LOAM_ODOM_BRUTE_FORCE": "1"}) == base
    assert _run({"LOAM_HOST_GN_LOOP": "1"}) == base
    assert _run({"LOAM_ODOM_BRUTE_FORCE": "1", "LOAM_HOST_GN_LOOP": "1"}) == base
