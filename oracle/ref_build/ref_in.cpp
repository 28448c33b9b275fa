// ORACLE — TEST INFRASTRUCTURE ONLY.  Compiles the reference's input_data.cpp unmodified and runs its main() to
// completion against in-memory "bags": every /velodyne_points publish is handed to the caller (who runs a SLAM pipeline or
// a stand-in and returns the odometry it produced); that odometry is delivered to the node's subscriber at its next
// ros::spinOnce(), as the ROS graph would; /control_command and /slam_track publishes are handed to the caller too.
#include <stdio.h>
#include <stdlib.h>
#include <fcntl.h>
#include <unistd.h>

#include "ref_common.h"
#define system(x) (0) /* the node clears the terminal every 50 messages */
#define main ref_in_main
#include "input_data.cpp"
#undef main
#undef system

struct RefReplayCallbacks {  // same layout as loam_replay_callbacks (include/loamgpu.h)
  void (*publish)(void* user, int bag_index, int msg_index, double* stamp, double odometry_xyzt[4], int* odometry_arrived);
  void (*control)(void* user);
  void (*slam_track)(void* user, int track_flag, const double* xyzt, int n_points);
  void* user;
};

REF_API int ref_in_run(const int* messages_per_bag, const double* stamps_flat, int n_bags, double long_distance, double short_distance,
                       double overlap_distance, const RefReplayCallbacks* cb, const char* baglist_path) {
  // bags: "refbag<k>", listed in a text file the node reads itself (IN:125-152)
  rosbag::registry().clear();
  FILE* f = fopen(baglist_path, "w");
  if (!f) return -1;
  int g = 0;
  for (int b = 0; b < n_bags; b++) {
    char name[64];
    snprintf(name, sizeof(name), "refbag%d", b);
    fprintf(f, "%s\n", name);
    std::vector<rosbag::MessageInstance>& v = rosbag::registry()[name];
    for (int m = 0; m < messages_per_bag[b]; m++, g++) {
      auto pc = std::make_shared<sensor_msgs::PointCloud2>();
      pc->header.stamp.fromSec(stamps_flat[g]);
      pc->header.frame_id = std::to_string(b) + " " + std::to_string(m);
      rosbag::MessageInstance mi;
      mi.msg = pc;
      v.push_back(mi);
    }
  }
  fclose(f);
  std::vector<std::shared_ptr<nav_msgs::Odometry>> pending;
  refshim::capture().on_publish = [&](const std::string& topic, const std::shared_ptr<void>& p) {
    if (topic == "/velodyne_points") {
      auto msg = *std::static_pointer_cast<sensor_msgs::PointCloud2::ConstPtr>(p);
      int b = 0, m = 0;
      sscanf(msg->header.frame_id.c_str(), "%d %d", &b, &m);
      double stamp = 0, odo[4] = {0, 0, 0, 0};
      int arrived = 0;
      cb->publish(cb->user, b, m, &stamp, odo, &arrived);
      if (arrived) {
        auto o = std::make_shared<nav_msgs::Odometry>();
        o->header.stamp.fromSec(odo[3]);
        o->pose.pose.position.x = odo[0]; o->pose.pose.position.y = odo[1]; o->pose.pose.position.z = odo[2];
        pending.push_back(o);
      }
    } else if (topic == "/control_command") {
      cb->control(cb->user);
    } else if (topic == "/slam_track") {
      auto t = std::static_pointer_cast<gpsCalibration::IMTrack>(p);
      std::vector<double> flat;
      for (auto& q : t->track) { flat.push_back(q.x); flat.push_back(q.y); flat.push_back(q.z); flat.push_back(q.t); }
      cb->slam_track(cb->user, (int)t->track_flag, flat.data(), (int)t->track.size());
    }
  };
  refshim::capture().on_spin_once = [&] {
    std::vector<std::shared_ptr<nav_msgs::Odometry>> batch;
    batch.swap(pending);
    auto it = refshim::subs().find("/true_odometry_to_init");
    for (auto& o : batch)
      if (it != refshim::subs().end()) it->second(o);
  };
  char a0[] = "input_data", a2[64], a3[64], a4[64], a5[] = "unused";
  snprintf(a2, sizeof(a2), "%.17g", long_distance);
  snprintf(a3, sizeof(a3), "%.17g", short_distance);
  snprintf(a4, sizeof(a4), "%.17g", overlap_distance);
  char* argv[] = {a0, (char*)baglist_path, a2, a3, a4, a5, nullptr};
  // the node chats on stdout: keep it away from the caller's
  fflush(stdout);
  int saved = dup(1), devnull = open("/dev/null", O_WRONLY);
  if (devnull >= 0) dup2(devnull, 1);
  int rc = ref_in_main(6, argv);
  fflush(stdout);
  std::cout.flush();
  if (saved >= 0) { dup2(saved, 1); close(saved); }
  if (devnull >= 0) close(devnull);
  refshim::capture().on_publish = nullptr;
  refshim::capture().on_spin_once = nullptr;
  return rc;
}
