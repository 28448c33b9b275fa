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
