// ORACLE — TEST INFRASTRUCTURE ONLY.  Compiles the reference's scanRegistration.cpp unmodified (from /root/reference)
// and exposes its laserCloudHandler (SR:238) through a small C API.
#include "ref_common.h"
#define main ref_node_main
#include "scanRegistration.cpp"
#undef main

static bool g_inited = false;
REF_API int ref_sr_process(const float* xyz, int n, double stamp) {
  if (!g_inited) {
    ref_node_main(0, nullptr);  // creates the publishers (ros::spin() returns immediately in the shim)
    g_inited = true;
  }
  auto m = refh::make_cloud(xyz, n, 3, stamp);
  laserCloudHandler(m);
  return 0;
}
// which: 0 /velodyne_cloud_2, 1 sharp, 2 less sharp, 3 flat, 4 less flat
REF_API int ref_sr_cloud(int which, float* buf, int cap) {
  const char* t[5] = {"/velodyne_cloud_2", "/laser_cloud_sharp", "/laser_cloud_less_sharp", "/laser_cloud_flat", "/laser_cloud_less_flat"};
  return refh::get_cloud(t[which], buf, cap, 4);
}

// One /imu/data message to the node's imuHandler (SR:754-837); the 12 floats of /imu_trans as published (SR:730-745).
REF_API int ref_sr_imu(double stamp, const double* q4, const double* av3, const double* la3) {
  auto m = boost::shared_ptr<sensor_msgs::Imu>(new sensor_msgs::Imu());
  m->header.stamp.fromSec(stamp);
  m->orientation.x = q4[0]; m->orientation.y = q4[1]; m->orientation.z = q4[2]; m->orientation.w = q4[3];
  m->angular_velocity.x = av3[0]; m->angular_velocity.y = av3[1]; m->angular_velocity.z = av3[2];
  m->linear_acceleration.x = la3[0]; m->linear_acceleration.y = la3[1]; m->linear_acceleration.z = la3[2];
  imuHandler(m);
  return 0;
}
REF_API int ref_sr_imu_trans(float* v12) { return refh::get_cloud("/imu_trans", v12, 4, 3); }
