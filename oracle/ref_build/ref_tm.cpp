// ORACLE — TEST INFRASTRUCTURE ONLY.  Compiles the reference's transformMaintenance.cpp unmodified; its main() keeps the
// publisher / broadcaster in locals and then spins, so it runs in a thread parked inside ros::spin() while the two
// handlers (TM:262-315, TM:317-338) are called directly.
#include "ref_common.h"
#define main ref_node_main
#include "transformMaintenance.cpp"
#undef main

static std::thread g_thread;
static bool g_started = false;

REF_API int ref_tm_start() {
  if (g_started) return 0;
  g_started = true;
  refshim::loop().spin_blocks = true;
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
// /laser_odom_to_init message in; out6 = transformMapped, track4 = /true_odometry_to_init {x, y, z, stamp}
REF_API int ref_tm_odometry(const double* pose7, double stamp, float* out6, double* track4) {
  laserOdometryHandler(make_odom(pose7, stamp));
  for (int i = 0; i < 6; i++) out6[i] = transformMapped[i];
  auto it = refshim::capture().last.find("/true_odometry_to_init");
  auto m = std::static_pointer_cast<nav_msgs::Odometry>(it->second);
  track4[0] = m->pose.pose.position.x; track4[1] = m->pose.pose.position.y; track4[2] = m->pose.pose.position.z;
  track4[3] = m->header.stamp.toSec();
  return 0;
}
// /aft_mapped_to_init message in: pose (position + quaternion) and the twist fields carrying transformBefMapped
REF_API int ref_tm_aft_mapped(const double* pose7, const float* bef6, double stamp) {
  auto m = make_odom(pose7, stamp);
  m->twist.twist.angular.x = bef6[0]; m->twist.twist.angular.y = bef6[1]; m->twist.twist.angular.z = bef6[2];
  m->twist.twist.linear.x = bef6[3]; m->twist.twist.linear.y = bef6[4]; m->twist.twist.linear.z = bef6[5];
  odomAftMappedHandler(m);
  return 0;
}
REF_API int ref_tm_stop() {
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
