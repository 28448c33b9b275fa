// ORACLE — TEST INFRASTRUCTURE ONLY.  Helpers shared by the three harnesses that wrap the reference's own nodes.
#pragma once
#include <thread>

#include "shim/refshim.h"

#define REF_API extern "C" __attribute__((visibility("default")))

namespace refh {
inline std::shared_ptr<sensor_msgs::PointCloud2> make_cloud(const float* p, int n, int fields, double stamp) {
  auto m = std::make_shared<sensor_msgs::PointCloud2>();
  m->fields = fields;
  m->data.assign(p, p + (size_t)n * fields);
  m->header.stamp.fromSec(stamp);
  return m;
}
inline int get_cloud(const std::string& topic, float* buf, int cap_points, int fields) {
  auto it = refshim::capture().last.find(topic);
  if (it == refshim::capture().last.end()) return 0;
  auto m = std::static_pointer_cast<sensor_msgs::PointCloud2>(it->second);
  int n = (int)(m->data.size() / m->fields);
  if (buf && cap_points >= n && m->fields == fields && n) std::memcpy(buf, m->data.data(), sizeof(float) * m->data.size());
  return n;
}
inline int pub_count(const std::string& topic) {
  auto it = refshim::capture().count.find(topic);
  return it == refshim::capture().count.end() ? 0 : it->second;
}
// deliver one message to the node's subscriber on `topic` (queued for the loop thread)
inline void post(const std::string& topic, std::shared_ptr<void> msg) {
  auto cb = refshim::subs().at(topic);
  refshim::loop().inbox.push_back([cb, msg] { cb(msg); });
}
// run one iteration of the node's `while (status)` loop with the posted messages
inline void run_loop_once() {
  refshim::Loop& L = refshim::loop();
  std::unique_lock<std::mutex> l(L.m);
  L.state = 1;
  L.cv.notify_all();
  L.cv.wait(l, [&] { return L.state == 0; });
}
inline void wait_parked() {
  refshim::Loop& L = refshim::loop();
  std::unique_lock<std::mutex> l(L.m);
  L.cv.wait(l, [&] { return L.state == 0; });
}
}  // namespace refh
