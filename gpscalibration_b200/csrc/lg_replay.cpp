// N1: the segment scheduler of input_data.cpp -- which messages of a bag list are (re)published to the SLAM
// pipeline, when the pipeline is reset, and which tracks go out on /slam_track -- without ROS or rosbag.  Host only.
// The reference interleaves this logic with rosbag iteration, sleeps and console output (IN:244-446) and keeps its
// state in file-scope globals updated by the odometry subscriber (IN:78-116); here the same decisions are taken by
// one function that is handed three callbacks (publish one message and return the odometry it produced, reset the
// pipeline, emit a track).  Two passes (IN:272): pass 0 cuts "long" tracks without overlap, pass 1 "short" tracks
// that overlap by `overlap_distance`; the travelled distance is measured on the SLAM output itself.
#include <math.h>

#include <deque>
#include <vector>

#include "../../include/loamgpu.h"

namespace {

struct Location {  // IN:57-63 DISTANCE
  int bag, msg;    // msg is 1-based like g_nMsgIndex
  double distance, timestamp;
};
struct Track {
  int flag;
  std::vector<double> xyzt;
};

struct Replay {
  const loam_replay_callbacks* cb;
  double slam_distance[2], overlap_distance[2];
  int times = 0;
  std::vector<Location> all_location;  // IN:65
  Location pub_location{0, 0, 0, 0};   // IN:66
  bool have_pre = false;               // preOdometry != NULL
  double pre[3] = {0, 0, 0};
  double total_distance = 0;  // g_dTotalDistane
  int bag_index = 0, msg_index = 0;
  Track slam_track{0, {}};
  std::deque<Track> track_queue;  // slamTrackVector
  loam_replay_stats stats{};

  void emit(const Track& t) {
    stats.tracks++;
    cb->slam_track(cb->user, t.flag, t.xyzt.data(), (int)(t.xyzt.size() / 4));
  }
  void control() {
    stats.resets++;
    cb->control(cb->user);
  }
  // publish one message, then what subOdometryHandler does with the odometry that comes back (IN:78-122)
  void publish_and_listen() {
    double stamp = 0, odo[4] = {0, 0, 0, 0};
    int arrived = 0;
    cb->publish(cb->user, bag_index, msg_index - 1, &stamp, odo, &arrived);
    stats.published++;
    if (!arrived) return;  // nothing came back (first sweep after a reset): the subscriber is simply not called
    if (odo[3] != stamp) {  // IN:80, IN:117-121
      stats.lost++;
      return;
    }
    slam_track.xyzt.insert(slam_track.xyzt.end(), odo, odo + 4);  // IN:82-87
    Location tmp{bag_index, msg_index, 0, odo[3]};
    if (have_pre) {  // IN:92-101
      tmp.distance = sqrt(pow(odo[0] - pre[0], 2) + pow(odo[1] - pre[1], 2) + pow(odo[2] - pre[2], 2)) + total_distance;
    }
    have_pre = true;
    pre[0] = odo[0]; pre[1] = odo[1]; pre[2] = odo[2];
    if (tmp.distance <= slam_distance[times] - overlap_distance[times]) {  // IN:105-108
      pub_location = tmp;
    } else if (all_location.back().timestamp != pub_location.timestamp) {  // IN:109-115
      all_location.push_back(pub_location);
    }
    total_distance = tmp.distance;
  }
};

}  // namespace

extern "C" int loam_replay_segments(const int* messages_per_bag, int n_bags, double long_distance, double short_distance,
                                    double overlap_distance, int first_pass, int last_pass, const loam_replay_callbacks* cb,
                                    loam_replay_stats* stats_out) {
  if (!messages_per_bag || n_bags < 0 || !cb || !cb->publish || !cb->control || !cb->slam_track) return LOAM_EINVAL;
  if (!(long_distance > short_distance && short_distance > overlap_distance && overlap_distance > 0)) return LOAM_EINVAL;  // IN:257
  if (first_pass < 0 || last_pass > 1 || first_pass > last_pass) return LOAM_EINVAL;
  for (int b = 0; b < n_bags; b++)
    if (messages_per_bag[b] < 0) return LOAM_EINVAL;
  Replay R;
  R.cb = cb;
  R.slam_distance[0] = long_distance; R.overlap_distance[0] = 0;  // IN:259-262
  R.slam_distance[1] = short_distance; R.overlap_distance[1] = overlap_distance;
  const double rest_divisor = 3.0;  // IMREST, IN:31
  for (R.times = first_pass; R.times <= last_pass; R.times++) {
    R.pub_location = Location{0, 0, 0, 0};  // IN:274-279
    R.all_location.push_back(R.pub_location);
    R.bag_index = 0;
    R.total_distance = 0;
    R.control();  // IN:281-285
    while (R.bag_index < n_bags) {  // IN:287
      bool end = false;
      R.bag_index = R.all_location.back().bag;  // IN:293
      while (R.bag_index < n_bags) {            // IN:304
        R.msg_index = 0;
        for (int m = 0; m < messages_per_bag[R.bag_index]; m++) {  // IN:313
          R.msg_index++;
          if (R.pub_location.msg < R.msg_index || R.pub_location.bag < R.bag_index) {  // IN:326
            R.publish_and_listen();                                                    // IN:328-335
            if (R.total_distance > R.slam_distance[R.times]) {                         // IN:336-344
              R.total_distance = 0;
              end = true;
              break;
            }
          }
        }
        if (end) {  // IN:348-353
          R.control();
          break;
        }
        ++R.bag_index;
      }
      R.slam_track.flag = R.times;  // IN:355-364
      R.track_queue.push_back(R.slam_track);
      R.slam_track.xyzt.clear();
      R.have_pre = false;
      if (R.track_queue.size() == 3) {
        R.emit(R.track_queue.front());
        R.track_queue.pop_front();
      }
    }
    // IN:367-424: a rest shorter than a third of a track is replayed once more, appended to the previous start
    if (R.all_location.size() > 1 && R.total_distance < R.slam_distance[R.times] / rest_divisor) {
      const Location tmp = R.all_location[R.all_location.size() - 2];
      R.track_queue.clear();
      R.slam_track.xyzt.clear();
      R.control();
      for (R.bag_index = tmp.bag; R.bag_index < n_bags; R.bag_index++) {
        R.msg_index = 0;
        for (int m = 0; m < messages_per_bag[R.bag_index]; m++) {
          R.msg_index++;
          if (tmp.msg < R.msg_index || tmp.bag < R.bag_index) R.publish_and_listen();
        }
      }
    }
    if (!R.slam_track.xyzt.empty()) {  // IN:428-433
      R.slam_track.flag = R.times;
      R.track_queue.push_back(R.slam_track);
      R.slam_track.xyzt.clear();
    }
    while (!R.track_queue.empty()) {  // IN:435-440
      R.emit(R.track_queue.front());
      R.track_queue.pop_front();
    }
    R.emit(R.slam_track);  // IN:441: the empty track that tells the consumers the pass is over
    R.all_location.clear();
  }
  if (stats_out) *stats_out = R.stats;
  return LOAM_OK;
}
