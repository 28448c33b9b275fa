"""ctypes binding of the synthetic sweep generator (csrc/synth.h).  Input data only."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

SENSOR_REF16 = 0   # 16 rings at the reference's ring-table angles (SR:303-318) x 1800 columns
SENSOR_VLP16 = 1   # true VLP-16 angles: exercises dropped beams / empty rings
SENSOR_HDL64 = 2   # 64 rings x 1875 columns


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "csrc", "libloamsynth.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        lib = C.CDLL(path)
        lib.loamsynth_create.restype = C.c_void_p
        lib.loamsynth_create.argtypes = [C.c_int, C.c_int, C.c_ulonglong]
        lib.loamsynth_destroy.argtypes = [C.c_void_p]
        lib.loamsynth_max_points.argtypes = [C.c_void_p]
        lib.loamsynth_rings.argtypes = [C.c_void_p]
        lib.loamsynth_sweep.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p]
        _LIB = lib
    return _LIB


class SweepGenerator:
    """Deterministic ray-cast sweeps: packed float32 xyz in the sensor frame, firing order."""

    def __init__(self, sensor=SENSOR_REF16, scene=0, seed=0xC0FFEE, t_offset=0.0):
        self._h = _lib().loamsynth_create(sensor, scene, seed)
        self.max_points = _lib().loamsynth_max_points(self._h)
        self.n_rings = _lib().loamsynth_rings(self._h)
        self.t_offset = float(t_offset)

    def sweep(self, sweep_id, out=None):
        """Returns (xyz[n,3] float32, pose6 float64 = x,y,z,yaw,pitch,roll of the sensor at sweep start)."""
        buf = np.empty((self.max_points, 3), np.float32) if out is None else out
        pose = np.zeros(6, np.float64)
        n = _lib().loamsynth_sweep(self._h, int(sweep_id), self.t_offset, buf.ctypes.data, pose.ctypes.data)
        return buf[:n], pose

    def __del__(self):
        try:
            _lib().loamsynth_destroy(self._h)
        except Exception:
            pass
