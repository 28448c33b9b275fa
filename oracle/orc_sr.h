// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_linalg.h header).
//
// CPU restatement of the reference's per-sweep feature extraction,
// src/gpsCalibration/src/lidar_slam/loam/scanRegistration.cpp (SR), laserCloudHandler SR:238-752, incl. the IMU branch
// (imuHandler SR:754-837, AccumulateIMUShift SR:187-233, the per-point de-skew SR:364-434 with SR:121-184; dormant in the
// shipped pipeline -- input_data only replays velodyne_points, so imuPointerLast stays -1 -- but live code).
// fp32 with the reference's expression order and its fp64 promotions (unsuffixed literals, M_PI, scanPeriod).
#pragma once
#include <cmath>
#include <cstring>
#include <vector>

#include "orc_cloud.h"

namespace orc {

struct SRParams {
  int n_scans = 16;          // SR:65
  int ring_mode = 0;         // 0 = the reference's 16-entry table SR:301-320; 1 = uniform table (HDL-64-shaped, outside the reference)
  float ring_ang_min = 0.f;  // ring_mode 1: ring = int((angle - ang_min)/ang_step + 0.5)
  float ring_ang_step = 1.f;
  double scan_period = 0.1;  // SR:56 (const double)
};

struct SRState {
  // The reference keeps these as file-scope arrays of POINTSNUM (SR:68-74) that persist between sweeps and are
  // only re-initialised on [5, cloudSize-5) (SR:454-478).  Capacity is lifted (CH:15 fence) but persistence is kept.
  std::vector<float> curvature;
  std::vector<int> sortInd, picked, label;
  void ensure(size_t n) {
    if (curvature.size() < n) {
      curvature.resize(n, 0.f);
      sortInd.resize(n, 0);
      picked.resize(n, 0);
      label.resize(n, 0);
    }
  }
};

struct SROut {
  Cloud full, sharp, lessSharp, flat, lessFlat;
  std::vector<int> scanStart, scanEnd;
  std::vector<int> pickedAfterMask;  // snapshot of cloudNeighborPicked after SR:492-549 (diagnostic)
};

// SR:297-320.  Returns -1 for "drop".
inline int ring_of(const SRParams& prm, float angle) {
  if (prm.ring_mode == 0) {
    int r = int(angle + (angle < 0.0 ? -0.5 : +0.5));
    switch (r) {
      case -15: return 0;
      case -13: return 1;
      case -11: return 2;
      case -9: return 3;
      case -7: return 4;
      case -5: return 5;
      case -4: return 6;
      case -3: return 7;
      case -2: return 8;
      case -1: return 9;
      case 0: return 10;
      case 1: return 11;
      case 3: return 12;
      case 5: return 13;
      case 7: return 14;
      case 9: return 15;
      default: return -1;
    }
  }
  float rel = (angle - prm.ring_ang_min) / prm.ring_ang_step;
  int r = int(rel + 0.5);
  if (rel + 0.5 < 0.0 || r >= prm.n_scans) return -1;
  return r;
}

inline float gap2(const P4& a, const P4& b) {
  float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
  return dx * dx + dy * dy + dz * dz;
}


// SR:597-622 / SR:641-666: mark up to 5 neighbours on each side until a gap^2 > 0.05 is met.
// FENCE (iv): the reference indexes ind+l without bounds; that is only in range when no ring is empty.  The
// guards below never fire for fully populated sweeps and turn the reference's out-of-bounds access into a stop.
inline void suppress_neighbours(const Cloud& c, int* picked, int ind, int cloudSize) {
  for (int l = 1; l <= 5; l++) {
    if (ind + l >= cloudSize || ind + l - 1 < 0) break;
    if (gap2(c[ind + l], c[ind + l - 1]) > 0.05) break;
    picked[ind + l] = 1;
  }
  for (int l = -1; l >= -5; l--) {
    if (ind + l < 0 || ind + l + 1 >= cloudSize) break;
    if (gap2(c[ind + l], c[ind + l + 1]) > 0.05) break;
    picked[ind + l] = 1;
  }
}

// ---- IMU state of the node (file-scope globals SR:76-110): a ring of 200 messages with integrated velocity / shift, the
// "Start" values latched at the sweep's first point, the "Cur" values interpolated per point -- all persistent between sweeps.
struct ImuState {
  static const int Q = 200;  // imuQueLength SR:78
  int front = 0, last = -1;  // imuPointerFront, imuPointerLast
  double time[Q] = {0};
  float roll[Q] = {0}, pitch[Q] = {0}, yaw[Q] = {0}, accX[Q] = {0}, accY[Q] = {0}, accZ[Q] = {0};
  float veloX[Q] = {0}, veloY[Q] = {0}, veloZ[Q] = {0}, shiftX[Q] = {0}, shiftY[Q] = {0}, shiftZ[Q] = {0};
  float rollStart = 0, pitchStart = 0, yawStart = 0, rollCur = 0, pitchCur = 0, yawCur = 0;
  float veloXStart = 0, veloYStart = 0, veloZStart = 0, shiftXStart = 0, shiftYStart = 0, shiftZStart = 0;
  float veloXCur = 0, veloYCur = 0, veloZCur = 0, shiftXCur = 0, shiftYCur = 0, shiftZCur = 0;
  float shiftFromStartXCur = 0, shiftFromStartYCur = 0, shiftFromStartZCur = 0;
  float veloFromStartXCur = 0, veloFromStartYCur = 0, veloFromStartZCur = 0;
  int imuMesg = 0;      // SR:59
  double initYaw = 0;   // SR:60
  // /imu_trans as published SR:730-745
  void trans12(float* v) const {
    v[0] = pitchStart; v[1] = yawStart; v[2] = rollStart; v[3] = pitchCur; v[4] = yawCur; v[5] = rollCur;
    v[6] = shiftFromStartXCur; v[7] = shiftFromStartYCur; v[8] = shiftFromStartZCur;
    v[9] = veloFromStartXCur; v[10] = veloFromStartYCur; v[11] = veloFromStartZCur;
  }
};

// tf::Matrix3x3(q).getRPY (double), as the shim and odometry_ros_hop state it
inline void quat_to_rpy(double qx, double qy, double qz, double qw, double& roll, double& pitch, double& yaw) {
  double d = qx * qx + qy * qy + qz * qz + qw * qw, s = 2.0 / d;
  double xs = qx * s, ys = qy * s, zs = qz * s;
  double wx = qw * xs, wy = qw * ys, wz = qw * zs, xx = qx * xs, xy = qx * ys, xz = qx * zs, yy = qy * ys, yz = qy * zs, zz = qz * zs;
  double m00 = 1.0 - (yy + zz), m01 = xy - wz, m02 = xz + wy, m10 = xy + wz, m20 = xz - wy, m21 = yz + wx, m22 = 1.0 - (xx + yy);
  if (std::fabs(m20) >= 1) {
    yaw = 0;
    if (m20 < 0) { pitch = M_PI / 2.0; roll = std::atan2(m01, m02); } else { pitch = -M_PI / 2.0; roll = std::atan2(-m01, -m02); }
  } else {
    pitch = -std::asin(m20);
    roll = std::atan2(m21 / std::cos(pitch), m22 / std::cos(pitch));
    yaw = std::atan2(m10 / std::cos(pitch), m00 / std::cos(pitch));
  }
}

// AccumulateIMUShift SR:187-233 (std::cos / std::sin of float arguments are the float overloads: SR:51-52)
inline void imu_accumulate(ImuState& s) {
  const int L = s.last;
  float roll = s.roll[L];
  float accX = s.accX[L], accY = s.accY[L], accZ = s.accZ[L];
  float x1 = cosf(roll) * accX - sinf(roll) * accY;
  float y1 = sinf(roll) * accX + cosf(roll) * accY;
  float z1 = accZ;
  accX = x1; accY = y1; accZ = z1;
  int back = (L + ImuState::Q - 1) % ImuState::Q;
  double timeDiff = s.time[L] - s.time[back];
  if (timeDiff < 0.2) {  // rfansScanPeriod SR:58
    s.shiftX[L] = (float)(s.shiftX[back] + s.veloX[back] * timeDiff + accX * timeDiff * timeDiff / 2);
    s.shiftY[L] = (float)(s.shiftY[back] + s.veloY[back] * timeDiff + accY * timeDiff * timeDiff / 2);
    s.shiftZ[L] = (float)(s.shiftZ[back] + s.veloZ[back] * timeDiff + accZ * timeDiff * timeDiff / 2);
    s.veloX[L] = (float)(s.veloX[back] + accX * timeDiff);
    s.veloY[L] = (float)(s.veloY[back] + accY * timeDiff);
    s.veloZ[L] = (float)(s.veloZ[back] + accZ * timeDiff);
  }
}

// imuHandler SR:754-837.  q = orientation {x, y, z, w}, av = angular velocity, la = linear acceleration.
inline void imu_handler(ImuState& s, double stamp, const double* q, const double* av, const double* la) {
  double roll, pitch, yaw;
  bool flag = false;
  s.imuMesg++;
  if (std::fabs(std::pow(q[0], 2) + std::pow(q[1], 2) + std::pow(q[2], 2) + std::pow(q[3], 2) - 1) < 0.1) {
    quat_to_rpy(q[0], q[1], q[2], q[3], roll, pitch, yaw);
    if (s.imuMesg == 1) s.initYaw = yaw;
  } else {
    return;
  }
  float accY = (float)(la[1] - std::sin(roll) * std::cos(pitch) * 9.81);
  float accZ = (float)(la[2] - std::cos(roll) * std::cos(pitch) * 9.81);
  float accX = (float)(la[0] + std::sin(pitch) * 9.81);
  s.last = (s.last + 1) % ImuState::Q;
  int back = (s.last + ImuState::Q - 1) % ImuState::Q;
  const double PI_CH = 3.141592653589;  // CH:17 #define PI
  if (s.imuMesg != 1) {
    if (av[2] > 3) {  // IMUANGULARNOISE CH:23
      if (s.yaw[back] > yaw) {
        if (std::fabs(s.yaw[back]) < PI_CH) { flag = true; yaw = s.yaw[back]; }
      }
    } else if (std::fabs(av[2]) < 3) {
      if (s.yaw[back] != yaw) { flag = true; yaw = s.yaw[back]; }
    } else if (av[2] < -1 * 3) {
      if (s.yaw[back] < yaw) {
        if (std::fabs(s.yaw[back]) < PI_CH) { flag = true; yaw = s.yaw[back]; }
      }
    }
  }
  s.time[s.last] = stamp;
  s.roll[s.last] = (float)roll;
  s.pitch[s.last] = (float)pitch;
  if (s.imuMesg != 1) {
    if (flag) s.yaw[s.last] = (float)yaw; else s.yaw[s.last] = (float)(yaw - s.initYaw);
  } else {
    s.yaw[s.last] = 0;
  }
  if (std::fabs(accX) > 2 || std::fabs(accY) > 2) return;
  s.accX[s.last] = accX; s.accY[s.last] = accY; s.accZ[s.last] = accZ;
  imu_accumulate(s);
}

// SR:364-434 for one kept point: advance imuPointerFront, interpolate the Cur values, latch the Start values at the sweep's
// first point (i == 0 in the NaN-filtered cloud) or de-skew the point (SR:121-184).
inline void imu_point(ImuState& s, double timeScanCur, float relTime, double scanPeriod, bool first_point, P4& point) {
  const int Q = ImuState::Q;
  float pointTime = (float)(relTime * scanPeriod);
  while (s.front != s.last) {
    if (timeScanCur + pointTime < s.time[s.front]) break;
    s.front = (s.front + 1) % Q;
  }
  const int F = s.front;
  if (timeScanCur + pointTime > s.time[F]) {
    if ((timeScanCur + pointTime) - s.time[F] < 0.2) {
      s.rollCur = s.roll[F]; s.pitchCur = s.pitch[F]; s.yawCur = s.yaw[F];
      s.veloXCur = s.veloX[F]; s.veloYCur = s.veloY[F]; s.veloZCur = s.veloZ[F];
      s.shiftXCur = s.shiftX[F]; s.shiftYCur = s.shiftY[F]; s.shiftZCur = s.shiftZ[F];
    }
  } else {
    if (s.time[F] - timeScanCur - pointTime < 0.2) {
      int B = (F + Q - 1) % Q;
      float ratioFront = (float)((timeScanCur + pointTime - s.time[B]) / (s.time[F] - s.time[B]));
      float ratioBack = (float)((s.time[F] - timeScanCur - pointTime) / (s.time[F] - s.time[B]));
      s.rollCur = s.roll[F] * ratioFront + s.roll[B] * ratioBack;
      s.pitchCur = s.pitch[F] * ratioFront + s.pitch[B] * ratioBack;
      if (s.yaw[F] - s.yaw[B] > M_PI) {
        s.yawCur = (float)(s.yaw[F] * ratioFront + (s.yaw[B] + 2 * M_PI) * ratioBack);
      } else if (s.yaw[F] - s.yaw[B] < -M_PI) {
        s.yawCur = (float)(s.yaw[F] * ratioFront + (s.yaw[B] - 2 * M_PI) * ratioBack);
      } else {
        s.yawCur = s.yaw[F] * ratioFront + s.yaw[B] * ratioBack;
      }
      s.veloXCur = s.veloX[F] * ratioFront + s.veloX[B] * ratioBack;
      s.veloYCur = s.veloY[F] * ratioFront + s.veloY[B] * ratioBack;
      s.veloZCur = s.veloZ[F] * ratioFront + s.veloZ[B] * ratioBack;
      s.shiftXCur = s.shiftX[F] * ratioFront + s.shiftX[B] * ratioBack;
      s.shiftYCur = s.shiftY[F] * ratioFront + s.shiftY[B] * ratioBack;
      s.shiftZCur = s.shiftZ[F] * ratioFront + s.shiftZ[B] * ratioBack;
    }
  }
  if (first_point) {
    s.rollStart = s.rollCur; s.pitchStart = s.pitchCur; s.yawStart = s.yawCur;
    s.veloXStart = s.veloXCur; s.veloYStart = s.veloYCur; s.veloZStart = s.veloZCur;
    s.shiftXStart = s.shiftXCur; s.shiftYStart = s.shiftYCur; s.shiftZStart = s.shiftZCur;
    return;
  }
  {  // ShiftToStartIMU SR:121-139
    s.shiftFromStartXCur = s.shiftXCur - s.shiftXStart - s.veloXStart * pointTime;
    s.shiftFromStartYCur = s.shiftYCur - s.shiftYStart - s.veloYStart * pointTime;
    s.shiftFromStartZCur = s.shiftZCur - s.shiftZStart - s.veloZStart * pointTime;
    float x1 = cosf(s.yawStart) * s.shiftFromStartXCur - sinf(s.yawStart) * s.shiftFromStartZCur;
    float y1 = s.shiftFromStartYCur;
    float z1 = sinf(s.yawStart) * s.shiftFromStartXCur + cosf(s.yawStart) * s.shiftFromStartZCur;
    float x2 = x1;
    float y2 = cosf(s.pitchStart) * y1 + sinf(s.pitchStart) * z1;
    float z2 = -sinf(s.pitchStart) * y1 + cosf(s.pitchStart) * z1;
    s.shiftFromStartXCur = cosf(s.rollStart) * x2 + sinf(s.rollStart) * y2;
    s.shiftFromStartYCur = -sinf(s.rollStart) * x2 + cosf(s.rollStart) * y2;
    s.shiftFromStartZCur = z2;
  }
  {  // VeloToStartIMU SR:142-160
    s.veloFromStartXCur = s.veloXCur - s.veloXStart;
    s.veloFromStartYCur = s.veloYCur - s.veloYStart;
    s.veloFromStartZCur = s.veloZCur - s.veloZStart;
    float x1 = cosf(s.yawStart) * s.veloFromStartXCur - sinf(s.yawStart) * s.veloFromStartZCur;
    float y1 = s.veloFromStartYCur;
    float z1 = sinf(s.yawStart) * s.veloFromStartXCur + cosf(s.yawStart) * s.veloFromStartZCur;
    float x2 = x1;
    float y2 = cosf(s.pitchStart) * y1 + sinf(s.pitchStart) * z1;
    float z2 = -sinf(s.pitchStart) * y1 + cosf(s.pitchStart) * z1;
    s.veloFromStartXCur = cosf(s.rollStart) * x2 + sinf(s.rollStart) * y2;
    s.veloFromStartYCur = -sinf(s.rollStart) * x2 + cosf(s.rollStart) * y2;
    s.veloFromStartZCur = z2;
  }
  {  // TransformToStartIMU SR:163-184
    float x1 = cosf(s.rollCur) * point.x - sinf(s.rollCur) * point.y;
    float y1 = sinf(s.rollCur) * point.x + cosf(s.rollCur) * point.y;
    float z1 = point.z;
    float x2 = x1;
    float y2 = cosf(s.pitchCur) * y1 - sinf(s.pitchCur) * z1;
    float z2 = sinf(s.pitchCur) * y1 + cosf(s.pitchCur) * z1;
    float x3 = cosf(s.yawCur) * x2 + sinf(s.yawCur) * z2;
    float y3 = y2;
    float z3 = -sinf(s.yawCur) * x2 + cosf(s.yawCur) * z2;
    float x4 = cosf(s.yawStart) * x3 - sinf(s.yawStart) * z3;
    float y4 = y3;
    float z4 = sinf(s.yawStart) * x3 + cosf(s.yawStart) * z3;
    float x5 = x4;
    float y5 = cosf(s.pitchStart) * y4 + sinf(s.pitchStart) * z4;
    float z5 = -sinf(s.pitchStart) * y4 + cosf(s.pitchStart) * z4;
    point.x = cosf(s.rollStart) * x5 + sinf(s.rollStart) * y5 + s.shiftFromStartXCur;
    point.y = -sinf(s.rollStart) * x5 + cosf(s.rollStart) * y5 + s.shiftFromStartYCur;
    point.z = z5 + s.shiftFromStartZCur;
  }
}

// xyz: n points, `stride` floats apart, sensor frame (x fwd, y left, z up).
// imu != nullptr with imu->last >= 0: the IMU branch SR:364-434 runs (timeScanCur = the sweep's stamp).
inline void extract(const SRParams& prm, SRState& st, const float* xyz, int n, int stride, SROut& out, ImuState* imu = nullptr,
                    double timeScanCur = 0.0) {
  const int R = prm.n_scans;
  out.full.clear(); out.sharp.clear(); out.lessSharp.clear(); out.flat.clear(); out.lessFlat.clear();
  out.scanStart.assign(R, 0);
  out.scanEnd.assign(R, 0);
  out.pickedAfterMask.clear();

  // SR:260-263 removeNaNFromPointCloud: keep points whose x, y, z are all finite, in order.
  std::vector<int> keep;
  keep.reserve(n);
  for (int i = 0; i < n; i++) {
    const float* p = xyz + (size_t)i * stride;
    if (std::isfinite(p[0]) && std::isfinite(p[1]) && std::isfinite(p[2])) keep.push_back(i);
  }
  int cloudSize = (int)keep.size();
  if (cloudSize == 0) return;
  const float* pf = xyz + (size_t)keep[0] * stride;
  const float* pl = xyz + (size_t)keep[cloudSize - 1] * stride;
  // SR:267-278
  float startOri = -atan2f(pf[1], pf[0]);
  float endOri = (float)(-atan2f(pl[1], pl[0]) + 2 * M_PI);
  if (endOri - startOri > 3 * M_PI) {
    endOri = (float)(endOri - 2 * M_PI);
  } else if (endOri - startOri < M_PI) {
    endOri = (float)(endOri + 2 * M_PI);
  }
  bool halfPassed = false;
  std::vector<Cloud> scans(R);
  // SR:284-437
  for (int t = 0; t < cloudSize; t++) {
    const float* p = xyz + (size_t)keep[t] * stride;
    P4 pt;
    pt.x = p[1];
    pt.y = p[2];
    pt.z = p[0];
    float angle = (float)(atanf(pt.y / sqrtf(pt.x * pt.x + pt.z * pt.z)) * 180 / M_PI);
    int scanID = ring_of(prm, angle);
    if (scanID < 0) continue;
    float ori = -atan2f(pt.x, pt.z);
    if (!halfPassed) {
      if (ori < startOri - M_PI / 2) {
        ori = (float)(ori + 2 * M_PI);
      } else if (ori > startOri + M_PI * 3 / 2) {
        ori = (float)(ori - 2 * M_PI);
      }
      if (ori - startOri > M_PI) halfPassed = true;
    } else {
      ori = (float)(ori + 2 * M_PI);
      if (ori < endOri - M_PI * 3 / 2) {
        ori = (float)(ori + 2 * M_PI);
      } else if (ori > endOri + M_PI / 2) {
        ori = (float)(ori - 2 * M_PI);
      }
    }
    float relTime = (ori - startOri) / (endOri - startOri);
    pt.i = (float)(scanID + prm.scan_period * relTime);
    if (imu && imu->last >= 0) imu_point(*imu, timeScanCur, relTime, prm.scan_period, t == 0, pt);  // SR:364-434
    scans[scanID].push_back(pt);
  }
  // SR:444-447
  Cloud& c = out.full;
  for (int r = 0; r < R; r++) c.insert(c.end(), scans[r].begin(), scans[r].end());
  cloudSize = (int)c.size();
  st.ensure((size_t)cloudSize + 16);
  float* curv = st.curvature.data();
  int* sortInd = st.sortInd.data();
  int* picked = st.picked.data();
  int* label = st.label.data();

  // SR:454-490
  int scanCount = -1;
  for (int i = 5; i < cloudSize - 5; i++) {
    float dX = c[i - 5].x + c[i - 4].x + c[i - 3].x + c[i - 2].x + c[i - 1].x - 10 * c[i].x + c[i + 1].x + c[i + 2].x +
               c[i + 3].x + c[i + 4].x + c[i + 5].x;
    float dY = c[i - 5].y + c[i - 4].y + c[i - 3].y + c[i - 2].y + c[i - 1].y - 10 * c[i].y + c[i + 1].y + c[i + 2].y +
               c[i + 3].y + c[i + 4].y + c[i + 5].y;
    float dZ = c[i - 5].z + c[i - 4].z + c[i - 3].z + c[i - 2].z + c[i - 1].z - 10 * c[i].z + c[i + 1].z + c[i + 2].z +
               c[i + 3].z + c[i + 4].z + c[i + 5].z;
    curv[i] = dX * dX + dY * dY + dZ * dZ;
    sortInd[i] = i;
    picked[i] = 0;
    label[i] = 0;
    if (int(c[i].i) != scanCount) {
      scanCount = int(c[i].i);
      if (scanCount > 0 && scanCount < R) {
        out.scanStart[scanCount] = i + 5;
        out.scanEnd[scanCount - 1] = i - 5;
      }
    }
  }
  out.scanStart[0] = 5;
  out.scanEnd[R - 1] = cloudSize - 5;

  // SR:492-549
  for (int i = 5; i < cloudSize - 6; i++) {
    float diff = gap2(c[i + 1], c[i]);
    if (diff > 0.1) {
      float depth1 = sqrtf(c[i].x * c[i].x + c[i].y * c[i].y + c[i].z * c[i].z);
      float depth2 = sqrtf(c[i + 1].x * c[i + 1].x + c[i + 1].y * c[i + 1].y + c[i + 1].z * c[i + 1].z);
      if (depth1 > depth2) {
        float dx = c[i + 1].x - c[i].x * depth2 / depth1;
        float dy = c[i + 1].y - c[i].y * depth2 / depth1;
        float dz = c[i + 1].z - c[i].z * depth2 / depth1;
        if (sqrtf(dx * dx + dy * dy + dz * dz) / depth2 < 0.1) {
          for (int l = 0; l <= 5; l++) picked[i - l] = 1;
        }
      } else {
        float dx = c[i + 1].x * depth1 / depth2 - c[i].x;
        float dy = c[i + 1].y * depth1 / depth2 - c[i].y;
        float dz = c[i + 1].z * depth1 / depth2 - c[i].z;
        if (sqrtf(dx * dx + dy * dy + dz * dz) / depth1 < 0.1) {
          for (int l = 1; l <= 6; l++) picked[i + l] = 1;
        }
      }
    }
    float diff2 = gap2(c[i], c[i - 1]);
    float dis = c[i].x * c[i].x + c[i].y * c[i].y + c[i].z * c[i].z;
    if (diff > 0.0002 * dis && diff2 > 0.0002 * dis) picked[i] = 1;
  }
  out.pickedAfterMask.assign(picked, picked + cloudSize);

  // SR:559-684
  for (int r = 0; r < R; r++) {
    Cloud lessFlatScan;
    for (int j = 0; j < 6; j++) {
      int sp = (out.scanStart[r] * (6 - j) + out.scanEnd[r] * j) / 6;
      int ep = (out.scanStart[r] * (5 - j) + out.scanEnd[r] * (j + 1)) / 6 - 1;
      // SR:568-576: stable ascending order by curvature (strict '<' bubble => equal keys keep their order)
      if (ep > sp)
        std::stable_sort(sortInd + sp, sortInd + ep + 1, [&](int a, int b) { return curv[a] < curv[b]; });

      int largest = 0;
      for (int k = ep; k >= sp; k--) {
        int ind = sortInd[k];
        if (picked[ind] == 0 && curv[ind] > 0.1) {
          largest++;
          if (largest <= 16) {
            label[ind] = 2;
            out.sharp.push_back(c[ind]);
            out.lessSharp.push_back(c[ind]);
          } else if (largest <= 20) {
            label[ind] = 1;
            out.lessSharp.push_back(c[ind]);
          } else {
            break;
          }
          picked[ind] = 1;
          suppress_neighbours(c, picked, ind, cloudSize);
        }
      }
      int smallest = 0;
      for (int k = sp; k <= ep; k++) {
        int ind = sortInd[k];
        if (picked[ind] == 0 && curv[ind] < 0.1) {
          label[ind] = -1;
          out.flat.push_back(c[ind]);
          smallest++;
          if (smallest >= 32) break;
          picked[ind] = 1;
          suppress_neighbours(c, picked, ind, cloudSize);
        }
      }
      for (int k = sp; k <= ep; k++)
        if (label[k] <= 0) lessFlatScan.push_back(c[k]);
    }
    Cloud ds;
    voxel_grid(lessFlatScan, 0.2f, ds);  // SR:677-683
    out.lessFlat.insert(out.lessFlat.end(), ds.begin(), ds.end());
  }
}

}  // namespace orc
