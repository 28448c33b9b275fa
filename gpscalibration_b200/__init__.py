"""gpscalibration_b200 — B200-native LiDAR registration hot path of gpsCalibration.

Only what the hot path needs lives here: ``csrc/`` holds the hand-written sm_100a kernels and the C ABI
(``include/loamgpu.h``); ``capi`` binds that ABI with ctypes; ``nodes`` mirrors the reference's three LOAM nodes
(scanRegistration / laserOdometry / laserMapping handler interfaces) on top of it; ``synth`` is the synthetic sweep
generator used by tests and bench.  There is no CPU fallback: without libloamgpu.so or a CUDA device, calls fail.
"""
from .capi import LoamGpu, LoamGpuPipeline, LoamError, load_library, library_path  # noqa: F401
from .nodes import ScanRegistration, LaserOdometry, LaserMapping, TransformMaintenance, LoamPipeline  # noqa: F401
from .synth import SweepGenerator  # noqa: F401
from .scheduler import SegmentScheduler, GpuSlam, replay_segments  # noqa: F401
