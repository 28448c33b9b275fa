// ORACLE — TEST INFRASTRUCTURE ONLY.  Compiles the reference's laserMapping.cpp unmodified and drives its main loop
// (LM:422-1144) one message set at a time.
#include "ref_common.h"
#define main ref_node_main
#include "laserMapping.cpp"
#undef main

static std::thread g_thread;
static bool g_started = false;

REF_API int ref_lm_start() {
  if (g_started) return 0;
  g_started = true;
  g_thread = std::thread([] { ref_node_main(0, nullptr); });
  refh::wait_parked();
  return 0;
}
static std::shared_ptr<nav_msgs::Odometry> make_odom(const double* pose7, double stamp) {
  auto m = std::make_shared<nav_msgs::Odometry>();
  m->header.stamp.fromSec(stamp);
  m->pose.pose.position.x = pose7[0]; m->pose.pose.position.y = pose7[1]; m->pose.pose.position.z = pose7[2];
  m->pose.pose.orientation.x = pose7[3]; m->pose.pose.orientation.y = pose7[4]; m->pose.pose.orientation.z = pose7[5];
  m->pose.pose.orientation.w = pose7[6];
  return m;
}
// An odometry message that arrives without clouds (every other sweep): only the handler runs (LM:314-335).
REF_API int ref_lm_odometry_only(const double* pose7, double stamp) {
  {
    std::lock_guard<std::mutex> l(refshim::loop().m);
    refh::post("/laser_odom_to_init", make_odom(pose7, stamp));
  }
  refh::run_loop_once();
  return 0;
}
// A full synchronised message set (LM:425-428).  out24 = transformAftMapped, transformBefMapped, transformTobeMapped,
// transformSum (6 each).
REF_API int ref_lm_step(const float* corner, int nc, const float* surf, int ns, const float* full, int nf, const double* pose7, double stamp,
                        float* out24) {
  {
    std::lock_guard<std::mutex> l(refshim::loop().m);
    refh::post("/laser_cloud_corner_last", refh::make_cloud(corner, nc, 4, stamp));
    refh::post("/laser_cloud_surf_last", refh::make_cloud(surf, ns, 4, stamp));
    refh::post("/velodyne_cloud_3", refh::make_cloud(full, nf, 4, stamp));
    refh::post("/laser_odom_to_init", make_odom(pose7, stamp));
  }
  refh::run_loop_once();
  for (int i = 0; i < 6; i++) {
    out24[i] = transformAftMapped[i];
    out24[6 + i] = transformBefMapped[i];
    out24[12 + i] = transformTobeMapped[i];
    out24[18 + i] = transformSum[i];
  }
  return 0;
}
// which: 0 /laser_cloud_surround, 1 /velodyne_cloud_registered
REF_API int ref_lm_cloud(int which, float* buf, int cap) {
  const char* t[2] = {"/laser_cloud_surround", "/velodyne_cloud_registered"};
  return refh::get_cloud(t[which], buf, cap, 4);
}
REF_API int ref_lm_map_size(int* n_corner, int* n_surf) {
  long a = 0, b = 0;
  for (int i = 0; i < laserCloudNum; i++) {
    a += (long)laserCloudCornerArray[i]->points.size();
    b += (long)laserCloudSurfArray[i]->points.size();
  }
  *n_corner = (int)a;
  *n_surf = (int)b;
  return 0;
}
REF_API int ref_lm_stop() {
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
