// Host-side scalar pose bookkeeping the reference keeps on the CPU (north_star: "the 6-DoF solve ... stay in host
// C++"): AccumulateRotation LO:292-309, PluginIMURotation LO:229-287, transformAssociateToMap LM:120-205 and the
// scalar pointAssociateToMap LM:244-262 used for pointOnYAxis (LM:483-487).  fp32, libm sinf/cosf, no FMA contraction
// (compile with -ffp-contract=off / nvcc -fmad=false -Xcompiler -ffp-contract=off).
#pragma once
#include <math.h>

namespace lgh {

// LO:229-287
static inline void plugin_imu_rotation(float bcx, float bcy, float bcz, float blx, float bly, float blz, float alx, float aly,
                                float alz, float& acx, float& acy, float& acz) {
  float sbcx = sinf(bcx), cbcx = cosf(bcx), sbcy = sinf(bcy), cbcy = cosf(bcy), sbcz = sinf(bcz), cbcz = cosf(bcz);
  float sblx = sinf(blx), cblx = cosf(blx), sbly = sinf(bly), cbly = cosf(bly), sblz = sinf(blz), cblz = cosf(blz);
  float salx = sinf(alx), calx = cosf(alx), saly = sinf(aly), caly = cosf(aly), salz = sinf(alz), calz = cosf(alz);
  float srx = -sbcx * (salx * sblx + calx * caly * cblx * cbly + calx * cblx * saly * sbly) -
              cbcx * cbcz * (calx * saly * (cbly * sblz - cblz * sblx * sbly) - calx * caly * (sbly * sblz + cbly * cblz * sblx) + cblx * cblz * salx) -
              cbcx * sbcz * (calx * caly * (cblz * sbly - cbly * sblx * sblz) - calx * saly * (cbly * cblz + sblx * sbly * sblz) + cblx * salx * sblz);
  acx = -asinf(srx);
  float srycrx = (cbcy * sbcz - cbcz * sbcx * sbcy) * (calx * saly * (cbly * sblz - cblz * sblx * sbly) - calx * caly * (sbly * sblz + cbly * cblz * sblx) + cblx * cblz * salx) -
                 (cbcy * cbcz + sbcx * sbcy * sbcz) * (calx * caly * (cblz * sbly - cbly * sblx * sblz) - calx * saly * (cbly * cblz + sblx * sbly * sblz) + cblx * salx * sblz) +
                 cbcx * sbcy * (salx * sblx + calx * caly * cblx * cbly + calx * cblx * saly * sbly);
  float crycrx = (cbcz * sbcy - cbcy * sbcx * sbcz) * (calx * caly * (cblz * sbly - cbly * sblx * sblz) - calx * saly * (cbly * cblz + sblx * sbly * sblz) + cblx * salx * sblz) -
                 (sbcy * sbcz + cbcy * cbcz * sbcx) * (calx * saly * (cbly * sblz - cblz * sblx * sbly) - calx * caly * (sbly * sblz + cbly * cblz * sblx) + cblx * cblz * salx) +
                 cbcx * cbcy * (salx * sblx + calx * caly * cblx * cbly + calx * cblx * saly * sbly);
  acy = atan2f(srycrx / cosf(acx), crycrx / cosf(acx));
  float srzcrx = sbcx * (cblx * cbly * (calz * saly - caly * salx * salz) - cblx * sbly * (caly * calz + salx * saly * salz) + calx * salz * sblx) -
                 cbcx * cbcz * ((caly * calz + salx * saly * salz) * (cbly * sblz - cblz * sblx * sbly) + (calz * saly - caly * salx * salz) * (sbly * sblz + cbly * cblz * sblx) - calx * cblx * cblz * salz) +
                 cbcx * sbcz * ((caly * calz + salx * saly * salz) * (cbly * cblz + sblx * sbly * sblz) + (calz * saly - caly * salx * salz) * (cblz * sbly - cbly * sblx * sblz) + calx * cblx * salz * sblz);
  float crzcrx = sbcx * (cblx * sbly * (caly * salz - calz * salx * saly) - cblx * cbly * (saly * salz + caly * calz * salx) + calx * calz * sblx) +
                 cbcx * cbcz * ((saly * salz + caly * calz * salx) * (sbly * sblz + cbly * cblz * sblx) + (caly * salz - calz * salx * saly) * (cbly * sblz - cblz * sblx * sbly) + calx * calz * cblx * cblz) -
                 cbcx * sbcz * ((saly * salz + caly * calz * salx) * (cblz * sbly - cbly * sblx * sblz) + (caly * salz - calz * salx * saly) * (cbly * cblz + sblx * sbly * sblz) - calx * calz * cblx * sblz);
  acz = atan2f(srzcrx / cosf(acx), crzcrx / cosf(acx));
}

// LO:292-309
static inline void accumulate_rotation(float cx, float cy, float cz, float lx, float ly, float lz, float& ox, float& oy, float& oz) {
  float srx = cosf(lx) * cosf(cx) * sinf(ly) * sinf(cz) - cosf(cx) * cosf(cz) * sinf(lx) - cosf(lx) * cosf(ly) * sinf(cx);
  ox = -asinf(srx);
  float srycrx = sinf(lx) * (cosf(cy) * sinf(cz) - cosf(cz) * sinf(cx) * sinf(cy)) +
                 cosf(lx) * sinf(ly) * (cosf(cy) * cosf(cz) + sinf(cx) * sinf(cy) * sinf(cz)) + cosf(lx) * cosf(ly) * cosf(cx) * sinf(cy);
  float crycrx = cosf(lx) * cosf(ly) * cosf(cx) * cosf(cy) - cosf(lx) * sinf(ly) * (cosf(cz) * sinf(cy) - cosf(cy) * sinf(cx) * sinf(cz)) -
                 sinf(lx) * (sinf(cy) * sinf(cz) + cosf(cy) * cosf(cz) * sinf(cx));
  oy = atan2f(srycrx / cosf(ox), crycrx / cosf(ox));
  float srzcrx = sinf(cx) * (cosf(lz) * sinf(ly) - cosf(ly) * sinf(lx) * sinf(lz)) +
                 cosf(cx) * sinf(cz) * (cosf(ly) * cosf(lz) + sinf(lx) * sinf(ly) * sinf(lz)) + cosf(lx) * cosf(cx) * cosf(cz) * sinf(lz);
  float crzcrx = cosf(lx) * cosf(lz) * cosf(cx) * cosf(cz) - cosf(cx) * sinf(cz) * (cosf(ly) * sinf(lz) - cosf(lz) * sinf(lx) * sinf(ly)) -
                 sinf(cx) * (sinf(ly) * sinf(lz) + cosf(ly) * cosf(lz) * sinf(lx));
  oz = atan2f(srzcrx / cosf(ox), crzcrx / cosf(ox));
}

// LM:120-205
static inline void transform_associate_to_map(const float* Tsum, const float* Tbef, const float* Taft, float* Tincre, float* Ttobe) {
  float x1 = cosf(Tsum[1]) * (Tbef[3] - Tsum[3]) - sinf(Tsum[1]) * (Tbef[5] - Tsum[5]);
  float y1 = Tbef[4] - Tsum[4];
  float z1 = sinf(Tsum[1]) * (Tbef[3] - Tsum[3]) + cosf(Tsum[1]) * (Tbef[5] - Tsum[5]);
  float x2 = x1;
  float y2 = cosf(Tsum[0]) * y1 + sinf(Tsum[0]) * z1;
  float z2 = -sinf(Tsum[0]) * y1 + cosf(Tsum[0]) * z1;
  Tincre[3] = cosf(Tsum[2]) * x2 + sinf(Tsum[2]) * y2;
  Tincre[4] = -sinf(Tsum[2]) * x2 + cosf(Tsum[2]) * y2;
  Tincre[5] = z2;

  float sbcx = sinf(Tsum[0]), cbcx = cosf(Tsum[0]), sbcy = sinf(Tsum[1]), cbcy = cosf(Tsum[1]), sbcz = sinf(Tsum[2]), cbcz = cosf(Tsum[2]);
  float sblx = sinf(Tbef[0]), cblx = cosf(Tbef[0]), sbly = sinf(Tbef[1]), cbly = cosf(Tbef[1]), sblz = sinf(Tbef[2]), cblz = cosf(Tbef[2]);
  float salx = sinf(Taft[0]), calx = cosf(Taft[0]), saly = sinf(Taft[1]), caly = cosf(Taft[1]), salz = sinf(Taft[2]), calz = cosf(Taft[2]);

  float srx = -sbcx * (salx * sblx + calx * cblx * salz * sblz + calx * calz * cblx * cblz) -
              cbcx * sbcy * (calx * calz * (cbly * sblz - cblz * sblx * sbly) - calx * salz * (cbly * cblz + sblx * sbly * sblz) + cblx * salx * sbly) -
              cbcx * cbcy * (calx * salz * (cblz * sbly - cbly * sblx * sblz) - calx * calz * (sbly * sblz + cbly * cblz * sblx) + cblx * cbly * salx);
  Ttobe[0] = -asinf(srx);

  float srycrx = sbcx * (cblx * cblz * (caly * salz - calz * salx * saly) - cblx * sblz * (caly * calz + salx * saly * salz) + calx * saly * sblx) -
                 cbcx * cbcy * ((caly * calz + salx * saly * salz) * (cblz * sbly - cbly * sblx * sblz) + (caly * salz - calz * salx * saly) * (sbly * sblz + cbly * cblz * sblx) - calx * cblx * cbly * saly) +
                 cbcx * sbcy * ((caly * calz + salx * saly * salz) * (cbly * cblz + sblx * sbly * sblz) + (caly * salz - calz * salx * saly) * (cbly * sblz - cblz * sblx * sbly) + calx * cblx * saly * sbly);
  float crycrx = sbcx * (cblx * sblz * (calz * saly - caly * salx * salz) - cblx * cblz * (saly * salz + caly * calz * salx) + calx * caly * sblx) +
                 cbcx * cbcy * ((saly * salz + caly * calz * salx) * (sbly * sblz + cbly * cblz * sblx) + (calz * saly - caly * salx * salz) * (cblz * sbly - cbly * sblx * sblz) + calx * caly * cblx * cbly) -
                 cbcx * sbcy * ((saly * salz + caly * calz * salx) * (cbly * sblz - cblz * sblx * sbly) + (calz * saly - caly * salx * salz) * (cbly * cblz + sblx * sbly * sblz) - calx * caly * cblx * sbly);
  Ttobe[1] = atan2f(srycrx / cosf(Ttobe[0]), crycrx / cosf(Ttobe[0]));

  float srzcrx = (cbcz * sbcy - cbcy * sbcx * sbcz) * (calx * salz * (cblz * sbly - cbly * sblx * sblz) - calx * calz * (sbly * sblz + cbly * cblz * sblx) + cblx * cbly * salx) -
                 (cbcy * cbcz + sbcx * sbcy * sbcz) * (calx * calz * (cbly * sblz - cblz * sblx * sbly) - calx * salz * (cbly * cblz + sblx * sbly * sblz) + cblx * salx * sbly) +
                 cbcx * sbcz * (salx * sblx + calx * cblx * salz * sblz + calx * calz * cblx * cblz);
  float crzcrx = (cbcy * sbcz - cbcz * sbcx * sbcy) * (calx * calz * (cbly * sblz - cblz * sblx * sbly) - calx * salz * (cbly * cblz + sblx * sbly * sblz) + cblx * salx * sbly) -
                 (sbcy * sbcz + cbcy * cbcz * sbcx) * (calx * salz * (cblz * sbly - cbly * sblx * sblz) - calx * calz * (sbly * sblz + cbly * cblz * sblx) + cblx * cbly * salx) +
                 cbcx * cbcz * (salx * sblx + calx * cblx * salz * sblz + calx * calz * cblx * cblz);
  Ttobe[2] = atan2f(srzcrx / cosf(Ttobe[0]), crzcrx / cosf(Ttobe[0]));

  x1 = cosf(Ttobe[2]) * Tincre[3] - sinf(Ttobe[2]) * Tincre[4];
  y1 = sinf(Ttobe[2]) * Tincre[3] + cosf(Ttobe[2]) * Tincre[4];
  z1 = Tincre[5];
  x2 = x1;
  y2 = cosf(Ttobe[0]) * y1 - sinf(Ttobe[0]) * z1;
  z2 = sinf(Ttobe[0]) * y1 + cosf(Ttobe[0]) * z1;
  Ttobe[3] = Taft[3] - (cosf(Ttobe[1]) * x2 + sinf(Ttobe[1]) * z2);
  Ttobe[4] = Taft[4] - y2;
  Ttobe[5] = Taft[5] - (-sinf(Ttobe[1]) * x2 + cosf(Ttobe[1]) * z2);
}


// LM:244-262 for a single host-side point (pointOnYAxis, LM:483-487)
static inline void associate_to_map(const float* T, const float* pi, float* po) {
  float x1 = cosf(T[2]) * pi[0] - sinf(T[2]) * pi[1];
  float y1 = sinf(T[2]) * pi[0] + cosf(T[2]) * pi[1];
  float z1 = pi[2];
  float x2 = x1;
  float y2 = cosf(T[0]) * y1 - sinf(T[0]) * z1;
  float z2 = sinf(T[0]) * y1 + cosf(T[0]) * z1;
  po[0] = cosf(T[1]) * x2 + sinf(T[1]) * z2 + T[3];
  po[1] = y2 + T[4];
  po[2] = -sinf(T[1]) * x2 + cosf(T[1]) * z2 + T[5];
}


// LO:1066-1078 -> LM:322-332 / TM:277-283: Euler angles -> quaternion message -> Euler angles, in double.
// createQuaternionMsgFromRollPitchYaw(rz, -rx, -ry) gives g; the message carries (-g.y, -g.z, g.x, g.w); the handler
// reads it as tf::Quaternion(o.z, -o.x, -o.y, o.w) = g again, builds the rotation matrix and calls getRPY.
inline void pose_message_hop(const float* in, float* out) {
  const double roll = in[2], pitch = -(double)in[0], yaw = -(double)in[1];
  const double cr = std::cos(roll * 0.5), sr = std::sin(roll * 0.5);
  const double cp = std::cos(pitch * 0.5), sp = std::sin(pitch * 0.5);
  const double cy = std::cos(yaw * 0.5), sy = std::sin(yaw * 0.5);
  const double gx = sr * cp * cy - cr * sp * sy;
  const double gy = cr * sp * cy + sr * cp * sy;
  const double gz = cr * cp * sy - sr * sp * cy;
  const double gw = cr * cp * cy + sr * sp * sy;
  const double msg[4] = {-gy, -gz, gx, gw};                       // geometry_msgs orientation {x, y, z, w}
  const double q[4] = {msg[2], -msg[0], -msg[1], msg[3]};         // the handler's tf::Quaternion {x, y, z, w}
  const double s2 = 2.0 / (q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  const double xs = q[0] * s2, ys = q[1] * s2, zs = q[2] * s2;
  const double wx = q[3] * xs, wy = q[3] * ys, wz = q[3] * zs;
  const double xx = q[0] * xs, xy = q[0] * ys, xz = q[0] * zs, yy = q[1] * ys, yz = q[1] * zs, zz = q[2] * zs;
  // the matrix entries getRPY reads
  const double r00 = 1.0 - (yy + zz), r01 = xy - wz, r02 = xz + wy, r10 = xy + wz, r20 = xz - wy, r21 = yz + wx, r22 = 1.0 - (xx + yy);
  double e_roll, e_pitch, e_yaw;
  if (std::fabs(r20) >= 1) {  // gimbal lock branch of tf::Matrix3x3::getEulerYPR
    e_yaw = 0;
    if (r20 < 0) {
      e_pitch = M_PI / 2.0;
      e_roll = std::atan2(r01, r02);
    } else {
      e_pitch = -M_PI / 2.0;
      e_roll = std::atan2(-r01, -r02);
    }
  } else {
    e_pitch = -std::asin(r20);
    e_roll = std::atan2(r21 / std::cos(e_pitch), r22 / std::cos(e_pitch));
    e_yaw = std::atan2(r10 / std::cos(e_pitch), r00 / std::cos(e_pitch));
  }
  out[0] = (float)-e_pitch;
  out[1] = (float)-e_yaw;
  out[2] = (float)e_roll;
  for (int i = 3; i < 6; i++) out[i] = (float)(double)in[i];  // position rides in float64 fields
}

// ---- the IMU side of scanRegistration that stays on the host: imuHandler (SR:754-837) and AccumulateIMUShift (SR:187-233).
// One message -> one ring entry with the integrated velocity / shift; the per-point de-skew (SR:364-434) is a kernel.
// R is lg_extract.h's SrImuRing (same layout on the device).
struct ImuHost {
  int last = -1;        // imuPointerLast
  int imuMesg = 0;      // SR:59
  double initYaw = 0;   // SR:60
  float accX[200] = {0}, accY[200] = {0}, accZ[200] = {0};
};
template <class Ring>
inline void imu_handler(ImuHost& s, Ring& R, double stamp, const double* q, const double* av, const double* la) {
  const int Q = 200;
  double roll, pitch, yaw;
  bool flag = false;
  s.imuMesg++;
  if (!(std::fabs(std::pow(q[0], 2) + std::pow(q[1], 2) + std::pow(q[2], 2) + std::pow(q[3], 2) - 1) < 0.1)) return;  // SR:760-768
  {  // tf::Matrix3x3(orientation).getRPY
    const double d = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3], sc = 2.0 / d;
    const double xs = q[0] * sc, ys = q[1] * sc, zs = q[2] * sc;
    const double wx = q[3] * xs, wy = q[3] * ys, wz = q[3] * zs, xx = q[0] * xs, xy = q[0] * ys, xz = q[0] * zs, yy = q[1] * ys, yz = q[1] * zs,
                 zz = q[2] * zs;
    const double m00 = 1.0 - (yy + zz), m01 = xy - wz, m02 = xz + wy, m10 = xy + wz, m20 = xz - wy, m21 = yz + wx, m22 = 1.0 - (xx + yy);
    if (std::fabs(m20) >= 1) {
      yaw = 0;
      if (m20 < 0) { pitch = M_PI / 2.0; roll = std::atan2(m01, m02); } else { pitch = -M_PI / 2.0; roll = std::atan2(-m01, -m02); }
    } else {
      pitch = -std::asin(m20);
      roll = std::atan2(m21 / std::cos(pitch), m22 / std::cos(pitch));
      yaw = std::atan2(m10 / std::cos(pitch), m00 / std::cos(pitch));
    }
  }
  if (s.imuMesg == 1) s.initYaw = yaw;
  const float accY = (float)(la[1] - std::sin(roll) * std::cos(pitch) * 9.81);  // SR:770-772
  const float accZ = (float)(la[2] - std::cos(roll) * std::cos(pitch) * 9.81);
  const float accX = (float)(la[0] + std::sin(pitch) * 9.81);
  s.last = (s.last + 1) % Q;
  const int L = s.last, back = (L + Q - 1) % Q;
  if (s.imuMesg != 1) {  // SR:777-809: the yaw may only move the way the angular velocity says
    const double PI_CH = 3.141592653589;  // CH:17
    if (av[2] > 3) {
      if (R.yaw[back] > yaw && std::fabs(R.yaw[back]) < PI_CH) { flag = true; yaw = R.yaw[back]; }
    } else if (std::fabs(av[2]) < 3) {
      if (R.yaw[back] != yaw) { flag = true; yaw = R.yaw[back]; }
    } else if (av[2] < -1 * 3) {
      if (R.yaw[back] < yaw && std::fabs(R.yaw[back]) < PI_CH) { flag = true; yaw = R.yaw[back]; }
    }
  }
  R.time[L] = stamp;
  R.roll[L] = (float)roll;
  R.pitch[L] = (float)pitch;
  if (s.imuMesg != 1) R.yaw[L] = flag ? (float)yaw : (float)(yaw - s.initYaw); else R.yaw[L] = 0;
  if (std::fabs(accX) > 2 || std::fabs(accY) > 2) return;  // SR:826-829: the entry keeps whatever the ring held there
  s.accX[L] = accX; s.accY[L] = accY; s.accZ[L] = accZ;
  // AccumulateIMUShift SR:187-233
  const float r = R.roll[L];
  const float ax = cosf(r) * accX - sinf(r) * accY, ay = sinf(r) * accX + cosf(r) * accY, az = accZ;
  const double timeDiff = R.time[L] - R.time[back];
  if (timeDiff < 0.2) {
    R.shiftX[L] = (float)(R.shiftX[back] + R.veloX[back] * timeDiff + ax * timeDiff * timeDiff / 2);
    R.shiftY[L] = (float)(R.shiftY[back] + R.veloY[back] * timeDiff + ay * timeDiff * timeDiff / 2);
    R.shiftZ[L] = (float)(R.shiftZ[back] + R.veloZ[back] * timeDiff + az * timeDiff * timeDiff / 2);
    R.veloX[L] = (float)(R.veloX[back] + ax * timeDiff);
    R.veloY[L] = (float)(R.veloY[back] + ay * timeDiff);
    R.veloZ[L] = (float)(R.veloZ[back] + az * timeDiff);
  }
}

}  // namespace lgh
