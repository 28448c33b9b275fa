"""Host-side mirror of the reference's three LOAM nodes on top of the C ABI.

The reference's interface for this path is the handler/loop-body of each ROS node:
  scanRegistration.cpp  laserCloudHandler(msg)          SR:238   -> ScanRegistration.laserCloudHandler(xyz, stamp)
  laserOdometry.cpp     main-loop body + controlHandler LO:498   -> LaserOdometry.spinOnce() / controlHandler(inited)
  laserMapping.cpp      main-loop body + odom handler   LM:422   -> LaserMapping.laserOdometryHandler(pose) / spinOnce()
Names, argument meaning and the "nothing published" behaviours follow the reference so the parity tests read like
tests of the reference's nodes.  ROS itself is absent here; messages are numpy arrays.
"""
import numpy as np

from .capi import LoamGpu


class ScanRegistration:
    """/velodyne_points -> /velodyne_cloud_2 + 4 feature clouds (SR:847-870)."""

    def __init__(self, gpu: LoamGpu):
        self.gpu = gpu
        self.counts = None

    def laserCloudHandler(self, xyz, stamp=0.0):
        self.counts = self.gpu.extract(xyz, stamp)
        return self.counts

    # published topics (materialised on demand; they stay device-resident otherwise)
    def velodyne_cloud_2(self):
        return self.gpu.cloud("full")

    def laser_cloud_sharp(self):
        return self.gpu.cloud("sharp")

    def laser_cloud_less_sharp(self):
        return self.gpu.cloud("less_sharp")

    def laser_cloud_flat(self):
        return self.gpu.cloud("flat")

    def laser_cloud_less_flat(self):
        return self.gpu.cloud("less_flat")


class LaserOdometry:
    """6 feature topics -> /laser_odom_to_init, /laser_cloud_{corner,surf}_last, /velodyne_cloud_3 (LO:432-472)."""

    def __init__(self, gpu: LoamGpu):
        self.gpu = gpu
        self.result = None

    def controlHandler(self, systemInited: bool):  # LO:411-415
        if not systemInited:
            self.gpu.reset()

    def spinOnce(self):
        """One loop body LO:502-1147 for the message set of the last laserCloudHandler call."""
        self.result = self.gpu.odometry_process()
        return self.result

    def laser_odom_to_init(self):
        return None if not self.result.odom_published else np.array(self.result.transform_sum, np.float32)

    def laser_cloud_corner_last(self):
        return self.gpu.cloud("corner_last")

    def laser_cloud_surf_last(self):
        return self.gpu.cloud("surf_last")

    def velodyne_cloud_3(self):
        return self.gpu.cloud("full_res3")


class LaserMapping:
    """corner/surf last + /velodyne_cloud_3 + odometry -> /aft_mapped_to_init, registered cloud, surround (LM:356-384)."""

    def __init__(self, gpu: LoamGpu):
        self.gpu = gpu
        self.result = None

    def laserOdometryHandler(self, transform_sum):  # LM:314-335
        self.gpu.mapping_odometry(transform_sum)

    def spinOnce(self):
        """One loop body LM:425-1139 (call when odometry published the full message set)."""
        self.result = self.gpu.mapping_process()
        return self.result

    def aft_mapped_to_init(self):
        return np.array(self.result.transform_aft_mapped, np.float32)

    def velodyne_cloud_registered(self):
        return self.gpu.cloud("registered")

    def laser_cloud_surround(self):
        return self.gpu.cloud("surround")


class TransformMaintenance:
    """/laser_odom_to_init + /aft_mapped_to_init -> /integrated_to_init, /true_odometry_to_init (TM:346-362)."""

    def __init__(self, gpu: LoamGpu):
        self.gpu = gpu

    def laserOdometryHandler(self, transform_sum, stamp):  # TM:262-315
        return self.gpu.integrate_odometry(transform_sum, stamp)

    def odomAftMappedHandler(self, aft_mapped, bef_mapped):  # TM:317-338
        self.gpu.integrate_mapping(aft_mapped, bef_mapped)


class LoamPipeline:
    """The three nodes wired the way gpsCalibration.launch wires them (LA:14-26), one sweep per call."""

    def __init__(self, device=0, **kw):
        self.gpu = LoamGpu(device=device, **kw)
        self.registration = ScanRegistration(self.gpu)
        self.odometry = LaserOdometry(self.gpu)
        self.mapping = LaserMapping(self.gpu)

    def reset(self):
        """IMControl{systemInited=false} (IN:281-284)."""
        self.odometry.controlHandler(False)

    def process(self, xyz, stamp=0.0):
        """Returns the C-ABI loam_sweep_result (one fused call: SR -> LO -> LM with device-resident hand-over)."""
        return self.gpu.process_sweep(xyz, stamp)

    def process_nodewise(self, xyz, stamp=0.0):
        """Same work through the three node mirrors (what three separate ROS processes would do)."""
        self.registration.laserCloudHandler(xyz, stamp)
        o = self.odometry.spinOnce()
        m = None
        if o.odom_published:
            self.mapping.laserOdometryHandler(o.transform_sum)
            if o.fullres_published:
                m = self.mapping.spinOnce()
        return o, m
