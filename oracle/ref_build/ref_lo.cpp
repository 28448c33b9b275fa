// ORACLE — TEST INFRASTRUCTURE ONLY.  Compiles the reference's laserOdometry.cpp unmodified and drives its main loop
// (LO:498-1151) one message set at a time.
#include "ref_common.h"
#define main ref_node_main
#include "laserOdometry.cpp"
#undef main

static std::thread g_thread;
static bool g_started = false;

REF_API int ref_lo_start() {
  if (g_started) return 0;
  g_started = true;
  g_thread = std::thread([] { ref_node_main(0, nullptr); });
  refh::wait_parked();
  return 0;
}
REF_API int ref_lo_control(int inited) {
  auto m = std::make_shared<gpsCalibration::IMControl>();
  m->systemInited = inited != 0;
  std::lock_guard<std::mutex> l(refshim::loop().m);
  refh::post("/control_command", m);
  return 0;
}
// One synchronised message set (LO:502-508).  out18 = transformationSum[6], transformation[6], then
// {odometry published, clouds published, full-res published, 0, 0, 0} for this step.
REF_API int ref_lo_step(const float* sharp, int ns, const float* less_sharp, int nls, const float* flat, int nf, const float* less_flat,
                        int nlf, const float* full, int nfull, const float* imu12, double stamp, float* out18) {
  int c_odom = refh::pub_count("/laser_odom_to_init"), c_cl = refh::pub_count("/laser_cloud_corner_last"),
      c_full = refh::pub_count("/velodyne_cloud_3");
  {
    std::lock_guard<std::mutex> l(refshim::loop().m);
    refh::post("/laser_cloud_sharp", refh::make_cloud(sharp, ns, 4, stamp));
    refh::post("/laser_cloud_less_sharp", refh::make_cloud(less_sharp, nls, 4, stamp));
    refh::post("/laser_cloud_flat", refh::make_cloud(flat, nf, 4, stamp));
    refh::post("/laser_cloud_less_flat", refh::make_cloud(less_flat, nlf, 4, stamp));
    refh::post("/velodyne_cloud_2", refh::make_cloud(full, nfull, 4, stamp));
    float zeros[12] = {0};
    refh::post("/imu_trans", refh::make_cloud(imu12 ? imu12 : zeros, 4, 3, stamp));
  }
  refh::run_loop_once();
  for (int i = 0; i < 6; i++) {
    out18[i] = transformationSum[i];
    out18[6 + i] = transformation[i];
  }
  out18[12] = (float)(refh::pub_count("/laser_odom_to_init") - c_odom);
  out18[13] = (float)(refh::pub_count("/laser_cloud_corner_last") - c_cl);
  out18[14] = (float)(refh::pub_count("/velodyne_cloud_3") - c_full);
  out18[15] = out18[16] = out18[17] = 0.f;
  return 0;
}
// the published /laser_odom_to_init message: position xyz + quaternion xyzw
REF_API int ref_lo_odometry(double* out7) {
  auto it = refshim::capture().last.find("/laser_odom_to_init");
  if (it == refshim::capture().last.end()) return 0;
  auto m = std::static_pointer_cast<nav_msgs::Odometry>(it->second);
  out7[0] = m->pose.pose.position.x; out7[1] = m->pose.pose.position.y; out7[2] = m->pose.pose.position.z;
  out7[3] = m->pose.pose.orientation.x; out7[4] = m->pose.pose.orientation.y; out7[5] = m->pose.pose.orientation.z;
  out7[6] = m->pose.pose.orientation.w;
  return 1;
}
// which: 0 /laser_cloud_corner_last, 1 /laser_cloud_surf_last, 2 /velodyne_cloud_3
REF_API int ref_lo_cloud(int which, float* buf, int cap) {
  const char* t[3] = {"/laser_cloud_corner_last", "/laser_cloud_surf_last", "/velodyne_cloud_3"};
  return refh::get_cloud(t[which], buf, cap, 4);
}
REF_API int ref_lo_stop() {
  if (!g_started) return 0;
  {
    std::lock_guard<std::mutex> l(refshim::loop().m);
    refshim::loop().stop = true;
    refshim::loop().cv.notify_all();
  }
  g_thread.join();
  g_started = false;
  return 0;
}
