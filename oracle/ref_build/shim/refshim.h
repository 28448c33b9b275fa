// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// Minimal stand-ins for the ROS / PCL / OpenCV / tf interfaces that the reference's three LOAM translation units
// (scanRegistration.cpp, laserOdometry.cpp, laserMapping.cpp) use, so that those files compile UNMODIFIED from
// /root/reference and run in-process (oracle/_ref/*.so).  Nothing here is copied from the third-party projects: the
// classes expose only the members the reference calls, and the numerical entry points forward to the oracle's
// restatements (orc_cloud.h: exact kNN, VoxelGrid; orc_linalg.h: GEMM, QR solve, Jacobi eigen, LU inverse).
// What this pins: every line of arithmetic and control flow that lives in the reference's own files.
// What it cannot pin: the third-party internals themselves (absent from /root/reference) — see oracle/README.md.
#pragma once
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../orc_cloud.h"
#include "../../orc_linalg.h"

namespace boost {
using std::shared_ptr;
}

// ------------------------------------------------------------------------------------------------ ROS core
namespace ros {
struct Time {
  double t = 0;
  Time() {}
  Time& fromSec(double s) {
    t = s;
    return *this;
  }
  double toSec() const { return t; }
};
}  // namespace ros

namespace std_msgs {
struct Header {
  ros::Time stamp;
  std::string frame_id;
};
}  // namespace std_msgs

namespace refshim {
// published messages, captured per topic (last message + how many were published)
struct Capture {
  std::map<std::string, std::shared_ptr<void>> last;
  std::map<std::string, int> count;
  // optional hooks for nodes that are driven synchronously (input_data): called on every publish / instead of the
  // thread handshake in spinOnce
  std::function<void(const std::string&, const std::shared_ptr<void>&)> on_publish;
  std::function<void()> on_spin_once;
};
inline Capture& capture() {
  static Capture c;
  return c;
}
// subscribers: topic -> callback taking a type-erased message
inline std::map<std::string, std::function<void(std::shared_ptr<void>)>>& subs() {
  static std::map<std::string, std::function<void(std::shared_ptr<void>)>> s;
  return s;
}
// main-loop handshake (laserOdometry / laserMapping run their `while (status)` loop in a thread)
struct Loop {
  std::mutex m;
  std::condition_variable cv;
  int state = 1;  // 1 = loop body running, 0 = parked in spinOnce waiting for the next message batch
  bool stop = false;
  bool spin_blocks = false;  // ros::spin() parks until stop (nodes whose main() keeps state in locals, e.g. transformMaintenance)
  std::vector<std::function<void()>> inbox;
};
inline Loop& loop() {
  static Loop l;
  return l;
}
}  // namespace refshim

namespace ros {
inline void init(int, char**, const char*) {}
inline bool ok() { return !refshim::loop().stop; }
inline void spin() {
  refshim::Loop& L = refshim::loop();
  if (!L.spin_blocks) return;
  std::unique_lock<std::mutex> l(L.m);
  L.state = 0;
  L.cv.notify_all();
  L.cv.wait(l, [&] { return L.stop; });
}
inline void spinOnce() {
  if (refshim::capture().on_spin_once) {
    refshim::capture().on_spin_once();
    return;
  }
  refshim::Loop& L = refshim::loop();
  std::vector<std::function<void()>> batch;
  {
    std::unique_lock<std::mutex> l(L.m);
    L.state = 0;
    L.cv.notify_all();
    L.cv.wait(l, [&] { return L.state == 1 || L.stop; });
    batch.swap(L.inbox);
  }
  for (auto& f : batch) f();
}
struct Rate {
  explicit Rate(double) {}
  void sleep() {}
};
struct Publisher {
  std::string topic;
  template <class M>
  void publish(const M& m) const {
    refshim::capture().last[topic] = std::make_shared<M>(m);
    refshim::capture().count[topic]++;
    if (refshim::capture().on_publish) refshim::capture().on_publish(topic, refshim::capture().last[topic]);
  }
};
struct Subscriber {};
struct NodeHandle {
  template <class M>
  Subscriber subscribe(const std::string& topic, int, void (*cb)(const boost::shared_ptr<M const>&)) {
    refshim::subs()[topic] = [cb](std::shared_ptr<void> p) { cb(std::static_pointer_cast<M const>(p)); };
    return Subscriber();
  }
  template <class M>
  Publisher advertise(const std::string& topic, int) {
    Publisher p;
    p.topic = topic;
    return p;
  }
};
}  // namespace ros
#define ROS_INFO(...) \
  do {                \
  } while (0)

// ------------------------------------------------------------------------------------------------ messages
namespace geometry_msgs {
struct Quaternion {
  double x = 0, y = 0, z = 0, w = 1;
};
struct Vector3 {
  double x = 0, y = 0, z = 0;
};
struct Point {
  double x = 0, y = 0, z = 0;
};
struct Pose {
  Point position;
  Quaternion orientation;
};
struct PoseWithCovariance {
  Pose pose;
};
struct Twist {
  Vector3 linear, angular;
};
struct TwistWithCovariance {
  Twist twist;
};
}  // namespace geometry_msgs

namespace sensor_msgs {
struct PointCloud2 {  // payload kept as floats: `fields` floats per point (3 = xyz, 4 = xyz + intensity)
  std_msgs::Header header;
  int fields = 4;
  std::vector<float> data;
  typedef boost::shared_ptr<PointCloud2 const> ConstPtr;
};
typedef boost::shared_ptr<PointCloud2 const> PointCloud2ConstPtr;
struct Imu {
  std_msgs::Header header;
  geometry_msgs::Quaternion orientation;
  geometry_msgs::Vector3 angular_velocity, linear_acceleration;
  typedef boost::shared_ptr<Imu const> ConstPtr;
};
}  // namespace sensor_msgs

namespace nav_msgs {
struct Odometry {
  std_msgs::Header header;
  std::string child_frame_id;
  geometry_msgs::PoseWithCovariance pose;
  geometry_msgs::TwistWithCovariance twist;
  typedef boost::shared_ptr<Odometry const> ConstPtr;
};
}  // namespace nav_msgs

namespace gpsCalibration {
struct IMLocalXYZT {
  double x = 0, y = 0, z = 0, t = 0;
};
struct IMLocalXYZTW {
  double x = 0, y = 0, z = 0, t = 0, w = 0;
};
struct IMTrack {
  std::vector<IMLocalXYZT> track;
  std::vector<IMLocalXYZTW> trackWithWeight;
  long long track_flag = 0;
  typedef boost::shared_ptr<IMTrack const> ConstPtr;
};
struct IMControl {
  bool systemInited = false;
  typedef boost::shared_ptr<IMControl const> ConstPtr;
};
}  // namespace gpsCalibration

// ------------------------------------------------------------------------------------------------ rosbag (input_data)
// A "bag file" is an in-memory list of PointCloud2 messages registered under its path before the node runs.
#define BOOST_FOREACH(decl, range) for (decl : range)
namespace rosbag {
struct BagIOException {};
namespace bagmode {
enum BagMode { Read = 1 };
}
struct MessageInstance {
  boost::shared_ptr<sensor_msgs::PointCloud2 const> msg;
  template <class M>
  boost::shared_ptr<M const> instantiate() const {
    return msg;
  }
};
inline std::map<std::string, std::vector<MessageInstance>>& registry() {
  static std::map<std::string, std::vector<MessageInstance>> r;
  return r;
}
struct Bag {
  const std::vector<MessageInstance>* cur = nullptr;
  void open(const std::string& path, int = bagmode::Read) {
    auto it = registry().find(path);
    if (it == registry().end()) throw BagIOException();
    cur = &it->second;
  }
  void close() { cur = nullptr; }
};
struct TopicQuery {
  explicit TopicQuery(const std::vector<std::string>&) {}
};
struct View {
  const std::vector<MessageInstance>* v;
  View(const Bag& b, const TopicQuery&) : v(b.cur) {}
  std::vector<MessageInstance>::const_iterator begin() const { return v->begin(); }
  std::vector<MessageInstance>::const_iterator end() const { return v->end(); }
};
}  // namespace rosbag

// ------------------------------------------------------------------------------------------------ PCL
namespace pcl {
struct PointXYZ {
  float x = 0, y = 0, z = 0;
};
struct PointXYZI {
  float x = 0, y = 0, z = 0, intensity = 0;
};
template <class T>
struct PointCloud {
  std::vector<T> points;
  unsigned width = 0, height = 1;
  bool is_dense = true;
  typedef boost::shared_ptr<PointCloud<T>> Ptr;
  PointCloud() {}
  PointCloud(unsigned w, unsigned h) : points((size_t)w * h), width(w), height(h) {}
  void push_back(const T& p) { points.push_back(p); }
  void clear() { points.clear(); }
  size_t size() const { return points.size(); }
  PointCloud& operator+=(const PointCloud& o) {
    points.insert(points.end(), o.points.begin(), o.points.end());
    return *this;
  }
};
inline void fromROSMsg(const sensor_msgs::PointCloud2& m, PointCloud<PointXYZ>& c) {
  size_t n = m.data.size() / m.fields;
  c.points.resize(n);
  for (size_t i = 0; i < n; i++) {
    c.points[i].x = m.data[i * m.fields + 0];
    c.points[i].y = m.data[i * m.fields + 1];
    c.points[i].z = m.data[i * m.fields + 2];
  }
}
inline void fromROSMsg(const sensor_msgs::PointCloud2& m, PointCloud<PointXYZI>& c) {
  size_t n = m.data.size() / m.fields;
  c.points.resize(n);
  for (size_t i = 0; i < n; i++) {
    c.points[i].x = m.data[i * m.fields + 0];
    c.points[i].y = m.data[i * m.fields + 1];
    c.points[i].z = m.data[i * m.fields + 2];
    c.points[i].intensity = m.fields > 3 ? m.data[i * m.fields + 3] : 0.f;
  }
}
inline void toROSMsg(const PointCloud<PointXYZ>& c, sensor_msgs::PointCloud2& m) {
  m.fields = 3;
  m.data.resize(c.points.size() * 3);
  for (size_t i = 0; i < c.points.size(); i++) {
    m.data[i * 3 + 0] = c.points[i].x;
    m.data[i * 3 + 1] = c.points[i].y;
    m.data[i * 3 + 2] = c.points[i].z;
  }
}
inline void toROSMsg(const PointCloud<PointXYZI>& c, sensor_msgs::PointCloud2& m) {
  m.fields = 4;
  m.data.resize(c.points.size() * 4);
  for (size_t i = 0; i < c.points.size(); i++) {
    m.data[i * 4 + 0] = c.points[i].x;
    m.data[i * 4 + 1] = c.points[i].y;
    m.data[i * 4 + 2] = c.points[i].z;
    m.data[i * 4 + 3] = c.points[i].intensity;
  }
}
template <class T>
inline void removeNaNFromPointCloud(const PointCloud<T>& in, PointCloud<T>& out, std::vector<int>& index) {
  std::vector<T> keep;
  index.clear();
  keep.reserve(in.points.size());
  for (size_t i = 0; i < in.points.size(); i++) {
    const T& p = in.points[i];
    if (std::isfinite(p.x) && std::isfinite(p.y) && std::isfinite(p.z)) {
      keep.push_back(p);
      index.push_back((int)i);
    }
  }
  out.points.swap(keep);
}
inline orc::Cloud to_orc(const PointCloud<PointXYZI>& c) {
  orc::Cloud o(c.points.size());
  for (size_t i = 0; i < o.size(); i++) o[i] = orc::P4{c.points[i].x, c.points[i].y, c.points[i].z, c.points[i].intensity};
  return o;
}
template <class T>
struct VoxelGrid {
  typename PointCloud<T>::Ptr input;
  float leaf = 0.1f;
  void setInputCloud(const typename PointCloud<T>::Ptr& c) { input = c; }
  void setLeafSize(float lx, float, float) { leaf = lx; }
  void filter(PointCloud<T>& out) {
    orc::Cloud in = to_orc(*input), res;
    orc::voxel_grid(in, leaf, res);
    out.points.resize(res.size());
    for (size_t i = 0; i < res.size(); i++) {
      out.points[i].x = res[i].x;
      out.points[i].y = res[i].y;
      out.points[i].z = res[i].z;
      out.points[i].intensity = res[i].i;
    }
  }
};
template <class T>
struct KdTreeFLANN {
  typedef boost::shared_ptr<KdTreeFLANN<T>> Ptr;
  orc::Cloud cloud;
  orc::KdTree tree;
  void setInputCloud(const typename PointCloud<T>::Ptr& c) {
    cloud = to_orc(*c);
    tree.build(cloud);
  }
  int nearestKSearch(const T& p, int k, std::vector<int>& idx, std::vector<float>& d2) const {
    std::vector<orc::Nbr> nb(k);
    int f = tree.knn(orc::P4{p.x, p.y, p.z, 0.f}, k, nb.data());
    idx.resize(f);
    d2.resize(f);
    for (int i = 0; i < f; i++) {
      idx[i] = nb[i].idx;
      d2[i] = nb[i].d2;
    }
    return f;
  }
};
}  // namespace pcl

// ------------------------------------------------------------------------------------------------ OpenCV
#define CV_32F 5
namespace cv {
enum { DECOMP_LU = 0, DECOMP_QR = 4 };
struct Scalar {
  double v;
  static Scalar all(double x) {
    Scalar s;
    s.v = x;
    return s;
  }
};
struct Mat {
  int rows = 0, cols = 0;
  std::vector<float> d;
  Mat() {}
  Mat(int r, int c, int, const Scalar& s) : rows(r), cols(c), d((size_t)r * c, (float)s.v) {}
  template <class F>
  F& at(int i, int j) {
    return d[(size_t)i * cols + j];
  }
  template <class F>
  const F& at(int i, int j) const {
    return d[(size_t)i * cols + j];
  }
  void copyTo(Mat& o) const { o = *this; }
  Mat inv() const {
    Mat o(rows, cols, CV_32F, Scalar::all(0));
    orc::lu_inverse(d.data(), o.d.data(), rows);
    return o;
  }
};
inline Mat operator*(const Mat& a, const Mat& b) {
  Mat c(a.rows, b.cols, CV_32F, Scalar::all(0));
  orc::gemm_f32_dacc(a.d.data(), b.d.data(), c.d.data(), a.rows, a.cols, b.cols);
  return c;
}
inline void transpose(const Mat& a, Mat& t) {
  Mat o(a.cols, a.rows, CV_32F, Scalar::all(0));
  for (int i = 0; i < a.rows; i++)
    for (int j = 0; j < a.cols; j++) o.d[(size_t)j * a.rows + i] = a.d[(size_t)i * a.cols + j];
  t = o;
}
inline bool solve(const Mat& A, const Mat& B, Mat& X, int) {
  if ((int)X.d.size() != A.cols) X = Mat(A.cols, 1, CV_32F, Scalar::all(0));
  return orc::qr_solve(A.d.data(), B.d.data(), X.d.data(), A.rows, A.cols);
}
inline bool eigen(const Mat& A, Mat& E, Mat& V) {
  int n = A.rows;
  if ((int)E.d.size() != n) E = Mat(1, n, CV_32F, Scalar::all(0));
  if ((int)V.d.size() != n * n) V = Mat(n, n, CV_32F, Scalar::all(0));
  orc::jacobi_eigen(A.d.data(), E.d.data(), V.d.data(), n);
  return true;
}
}  // namespace cv

// ------------------------------------------------------------------------------------------------ tf
namespace tf {
struct Vector3 {
  double x, y, z;
  Vector3(double a = 0, double b = 0, double c = 0) : x(a), y(b), z(c) {}
};
struct Quaternion {
  double x, y, z, w;
  Quaternion(double a = 0, double b = 0, double c = 0, double d = 1) : x(a), y(b), z(c), w(d) {}
};
inline geometry_msgs::Quaternion createQuaternionMsgFromRollPitchYaw(double roll, double pitch, double yaw) {
  double hy = yaw * 0.5, hp = pitch * 0.5, hr = roll * 0.5;
  double cy = std::cos(hy), sy = std::sin(hy), cp = std::cos(hp), sp = std::sin(hp), cr = std::cos(hr), sr = std::sin(hr);
  geometry_msgs::Quaternion q;
  q.x = sr * cp * cy - cr * sp * sy;
  q.y = cr * sp * cy + sr * cp * sy;
  q.z = cr * cp * sy - sr * sp * cy;
  q.w = cr * cp * cy + sr * sp * sy;
  return q;
}
inline void quaternionMsgToTF(const geometry_msgs::Quaternion& m, Quaternion& q) { q = Quaternion(m.x, m.y, m.z, m.w); }
struct Matrix3x3 {
  double m[3][3];
  explicit Matrix3x3(const Quaternion& q) {
    double d = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
    double s = 2.0 / d;
    double xs = q.x * s, ys = q.y * s, zs = q.z * s;
    double wx = q.w * xs, wy = q.w * ys, wz = q.w * zs;
    double xx = q.x * xs, xy = q.x * ys, xz = q.x * zs;
    double yy = q.y * ys, yz = q.y * zs, zz = q.z * zs;
    m[0][0] = 1.0 - (yy + zz); m[0][1] = xy - wz; m[0][2] = xz + wy;
    m[1][0] = xy + wz; m[1][1] = 1.0 - (xx + zz); m[1][2] = yz - wx;
    m[2][0] = xz - wy; m[2][1] = yz + wx; m[2][2] = 1.0 - (xx + yy);
  }
  void getRPY(double& roll, double& pitch, double& yaw) const {
    if (std::fabs(m[2][0]) >= 1) {
      yaw = 0;
      if (m[2][0] < 0) {
        pitch = M_PI / 2.0;
        roll = std::atan2(m[0][1], m[0][2]);
      } else {
        pitch = -M_PI / 2.0;
        roll = std::atan2(-m[0][1], -m[0][2]);
      }
    } else {
      pitch = -std::asin(m[2][0]);
      roll = std::atan2(m[2][1] / std::cos(pitch), m[2][2] / std::cos(pitch));
      yaw = std::atan2(m[1][0] / std::cos(pitch), m[0][0] / std::cos(pitch));
    }
  }
};
struct StampedTransform {
  ros::Time stamp_;
  std::string frame_id_, child_frame_id_;
  void setRotation(const Quaternion&) {}
  void setOrigin(const Vector3&) {}
};
struct TransformBroadcaster {
  void sendTransform(const StampedTransform&) {}
};
}  // namespace tf
