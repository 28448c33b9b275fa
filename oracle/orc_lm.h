// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_linalg.h header).
//
// CPU restatement of the reference's scan-to-map stage,
// src/gpsCalibration/src/lidar_slam/loam/laserMapping.cpp (LM): transformAssociateToMap LM:120-205, transformUpdate
// LM:207-242 (IMU branch dormant), pointAssociateToMap LM:244-262, pointAssociateTobeMapped LM:264-282 and the
// main-loop body LM:425-1139.  The odometry pose enters as transformSum[6] directly; the reference's
// quaternion -> RPY round trip in laserOdometryHandler (LM:322-332) belongs to the ROS adapter.
#pragma once
#include <cmath>
#include <cstring>
#include <vector>

#include "orc_cloud.h"
#include "orc_linalg.h"
#include "orc_lo.h"

namespace orc {

// LM:244-262
inline void associate_to_map(const float* T, const P4& pi, P4& po) {
  float x1 = cosf(T[2]) * pi.x - sinf(T[2]) * pi.y;
  float y1 = sinf(T[2]) * pi.x + cosf(T[2]) * pi.y;
  float z1 = pi.z;
  float x2 = x1;
  float y2 = cosf(T[0]) * y1 - sinf(T[0]) * z1;
  float z2 = sinf(T[0]) * y1 + cosf(T[0]) * z1;
  po.x = cosf(T[1]) * x2 + sinf(T[1]) * z2 + T[3];
  po.y = y2 + T[4];
  po.z = -sinf(T[1]) * x2 + cosf(T[1]) * z2 + T[5];
  po.i = pi.i;
}

// LM:264-282
inline void associate_tobe_mapped(const float* T, const P4& pi, P4& po) {
  float x1 = cosf(T[1]) * (pi.x - T[3]) - sinf(T[1]) * (pi.z - T[5]);
  float y1 = pi.y - T[4];
  float z1 = sinf(T[1]) * (pi.x - T[3]) + cosf(T[1]) * (pi.z - T[5]);
  float x2 = x1;
  float y2 = cosf(T[0]) * y1 + sinf(T[0]) * z1;
  float z2 = -sinf(T[0]) * y1 + cosf(T[0]) * z1;
  po.x = cosf(T[2]) * x2 + sinf(T[2]) * y2;
  po.y = -sinf(T[2]) * x2 + cosf(T[2]) * y2;
  po.z = z2;
  po.i = pi.i;
}

// LM:120-205
inline void transform_associate_to_map(const float* Tsum, const float* Tbef, const float* Taft, float* Tincre, float* Ttobe) {
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

// The ROS hop between the two nodes: laserOdometry publishes transformSum as a quaternion built from
// (rz, -rx, -ry) with the axes permuted (LO:1066-1078) and laserMapping's handler turns it back into
// (-pitch, -yaw, roll) (LM:322-332), all in double.  Identity up to an ulp; restated so the oracle can be compared
// bit-for-bit with the reference's own nodes wired through messages.
inline void odometry_ros_hop(const float* Tsum_in, float* Tsum_out) {
  double roll = Tsum_in[2], pitch = -Tsum_in[0], yaw = -Tsum_in[1];
  double hy = yaw * 0.5, hp = pitch * 0.5, hr = roll * 0.5;
  double cy = std::cos(hy), sy = std::sin(hy), cp = std::cos(hp), sp = std::sin(hp), cr = std::cos(hr), sr = std::sin(hr);
  double gx = sr * cp * cy - cr * sp * sy, gy = cr * sp * cy + sr * cp * sy, gz = cr * cp * sy - sr * sp * cy, gw = cr * cp * cy + sr * sp * sy;
  double ox = -gy, oy = -gz, oz = gx, ow = gw;   // published orientation
  double qx = oz, qy = -ox, qz = -oy, qw = ow;   // tf::Quaternion(geoQuat.z, -geoQuat.x, -geoQuat.y, geoQuat.w)
  double d = qx * qx + qy * qy + qz * qz + qw * qw, s = 2.0 / d;
  double xs = qx * s, ys = qy * s, zs = qz * s;
  double wx = qw * xs, wy = qw * ys, wz = qw * zs, xx = qx * xs, xy = qx * ys, xz = qx * zs, yy = qy * ys, yz = qy * zs, zz = qz * zs;
  double m00 = 1.0 - (yy + zz), m10 = xy + wz, m20 = xz - wy, m21 = yz + wx, m22 = 1.0 - (xx + yy), m01 = xy - wz, m02 = xz + wy;
  double r, p, y;
  if (std::fabs(m20) >= 1) {
    y = 0;
    if (m20 < 0) { p = M_PI / 2.0; r = std::atan2(m01, m02); } else { p = -M_PI / 2.0; r = std::atan2(-m01, -m02); }
  } else {
    p = -std::asin(m20);
    r = std::atan2(m21 / std::cos(p), m22 / std::cos(p));
    y = std::atan2(m10 / std::cos(p), m00 / std::cos(p));
  }
  Tsum_out[0] = (float)-p;
  Tsum_out[1] = (float)-y;
  Tsum_out[2] = (float)r;
  Tsum_out[3] = (float)(double)Tsum_in[3];
  Tsum_out[4] = (float)(double)Tsum_in[4];
  Tsum_out[5] = (float)(double)Tsum_in[5];
}

// transformMaintenance.cpp (TM): fuses the per-sweep odometry pose with the latest mapping correction
// (laserOdometryHandler TM:262-315, odomAftMappedHandler TM:317-338; its transformAssociateToMap TM:175-260 is the
// same function as LM:120-205) and builds the "height compensated" planar track /true_odometry_to_init
// (SaveTrailWithTimeTotxt TM:116-157).  SURVEY 8f row N2.
struct TransformMaintenance {
  float Tsum[6] = {0}, Tincre[6] = {0}, Tmapped[6] = {0}, Tbef[6] = {0}, Taft[6] = {0};
  double preX = 0, preY = 0, preZ = 0, preT = 0, tmpX = 0, tmpY = 0, tmpZ = 0, tmpT = 0;
  // out6 = /integrated_to_init pose (transformMapped); track4 = /true_odometry_to_init {x, y, HEIGHT = 10, stamp}
  void odometry(const float* Tsum_in, double stamp, float* out6, double* track4) {
    if (fabs((double)Tsum_in[3]) < 0.000001 && fabs((double)Tsum_in[4]) < 0.000001 && fabs((double)Tsum_in[5]) < 0.000001) {
      preT = 0;
      for (int i = 0; i < 6; i++) Tsum[i] = Tincre[i] = Tmapped[i] = Tbef[i] = Taft[i] = 0;
    }
    for (int i = 0; i < 6; i++) Tsum[i] = Tsum_in[i];
    transform_associate_to_map(Tsum, Tbef, Taft, Tincre, Tmapped);
    for (int i = 0; i < 6; i++) out6[i] = Tmapped[i];
    // TM:116-157; laserOdometry2.pose.position = (transformMapped[3], [4], [5]) as doubles
    double px = Tmapped[3], py = Tmapped[4], pz = Tmapped[5];
    if (preT == 0) {
      preX = pz; preY = px; preZ = py; preT = stamp;
      tmpX = preX; tmpY = preY; tmpZ = preZ; tmpT = preT;
    } else {
      double dX = pz - preX, dY = px - preY, dZ = py - preZ;
      double dX1 = dX * sqrt(pow(dX, 2) + pow(dY, 2) + pow(dZ, 2)) / sqrt(pow(dX, 2) + pow(dY, 2));
      double dY1 = dY * sqrt(pow(dX, 2) + pow(dY, 2) + pow(dZ, 2)) / sqrt(pow(dX, 2) + pow(dY, 2));
      tmpX += dX1; tmpY += dY1; tmpZ = py; tmpT = stamp;
      preX = pz; preY = px; preZ = py; preT = stamp;
    }
    track4[0] = tmpX; track4[1] = tmpY; track4[2] = 10; track4[3] = tmpT;
  }
  void aft_mapped(const float* aft6, const float* bef6) {
    for (int i = 0; i < 6; i++) { Taft[i] = aft6[i]; Tbef[i] = bef6[i]; }
  }
};

struct MapCorr {  // 5 neighbour indices per stack point, -1 when the 5th neighbour is not within 1 m (diagnostic)
  std::vector<int> corner, surf;
};

// One pass of the iteration body LM:754-964 (no solve).
inline void map_iteration(const Cloud& cornerStack, const Cloud& surfStack, const Cloud& cornerMap, const Cloud& surfMap,
                          const KnnIndex& kCorner, const KnnIndex& kSurf, const float* T, MapCorr* corr, NormalEq& ne,
                          int min_rows = 50) {
  std::vector<P4> ori, coef;
  P4 sel;
  Nbr nb[5];
  if (corr) {
    corr->corner.assign(cornerStack.size() * 5, -1);
    corr->surf.assign(surfStack.size() * 5, -1);
  }
  for (size_t i = 0; i < cornerStack.size(); i++) {
    const P4& po = cornerStack[i];
    associate_to_map(T, po, sel);
    int found = kCorner.knn(sel, 5, nb);
    if (found == 5 && nb[4].d2 < 1.0) {
      if (corr)
        for (int j = 0; j < 5; j++) corr->corner[i * 5 + j] = nb[j].idx;
      float cx = 0, cy = 0, cz = 0;
      for (int j = 0; j < 5; j++) {
        cx += cornerMap[nb[j].idx].x;
        cy += cornerMap[nb[j].idx].y;
        cz += cornerMap[nb[j].idx].z;
      }
      cx /= 5; cy /= 5; cz /= 5;
      float a11 = 0, a12 = 0, a13 = 0, a22 = 0, a23 = 0, a33 = 0;
      for (int j = 0; j < 5; j++) {
        float ax = cornerMap[nb[j].idx].x - cx;
        float ay = cornerMap[nb[j].idx].y - cy;
        float az = cornerMap[nb[j].idx].z - cz;
        a11 += ax * ax; a12 += ax * ay; a13 += ax * az;
        a22 += ay * ay; a23 += ay * az; a33 += az * az;
      }
      a11 /= 5; a12 /= 5; a13 /= 5; a22 /= 5; a23 /= 5; a33 /= 5;
      float A1[9] = {a11, a12, a13, a12, a22, a23, a13, a23, a33};
      float D1[3], V1[9];
      jacobi_eigen(A1, D1, V1, 3);  // LM:810
      if (D1[0] > 3 * D1[1]) {
        float x0 = sel.x, y0 = sel.y, z0 = sel.z;
        float x1 = (float)(cx + 0.1 * V1[0]), y1 = (float)(cy + 0.1 * V1[1]), z1 = (float)(cz + 0.1 * V1[2]);
        float x2 = (float)(cx - 0.1 * V1[0]), y2 = (float)(cy - 0.1 * V1[1]), z2 = (float)(cz - 0.1 * V1[2]);
        float la, lb, lc, ld2;
        line_coeff(x0, y0, z0, x1, y1, z1, x2, y2, z2, la, lb, lc, ld2);
        float s = (float)(1 - 0.9 * fabsf(ld2));
        if (s > 0.1) {
          ori.push_back(po);
          coef.push_back(P4{s * la, s * lb, s * lc, s * ld2});
        }
      }
    }
  }
  for (size_t i = 0; i < surfStack.size(); i++) {
    const P4& po = surfStack[i];
    associate_to_map(T, po, sel);
    int found = kSurf.knn(sel, 5, nb);
    if (found == 5 && nb[4].d2 < 1.0) {
      if (corr)
        for (int j = 0; j < 5; j++) corr->surf[i * 5 + j] = nb[j].idx;
      float A0[15], B0[5] = {-1, -1, -1, -1, -1}, X0[3];
      for (int j = 0; j < 5; j++) {
        A0[j * 3 + 0] = surfMap[nb[j].idx].x;
        A0[j * 3 + 1] = surfMap[nb[j].idx].y;
        A0[j * 3 + 2] = surfMap[nb[j].idx].z;
      }
      qr_solve(A0, B0, X0, 5, 3);  // LM:875
      float pa = X0[0], pb = X0[1], pc = X0[2], pd = 1;
      float ps = sqrtf(pa * pa + pb * pb + pc * pc);
      pa /= ps; pb /= ps; pc /= ps; pd /= ps;
      bool planeValid = true;
      for (int j = 0; j < 5; j++) {
        if (fabsf(pa * surfMap[nb[j].idx].x + pb * surfMap[nb[j].idx].y + pc * surfMap[nb[j].idx].z + pd) > 0.2) {
          planeValid = false;
          break;
        }
      }
      if (planeValid) {
        float pd2 = pa * sel.x + pb * sel.y + pc * sel.z + pd;
        float s = (float)(1 - 0.9 * fabsf(pd2) / sqrtf(sqrtf(sel.x * sel.x + sel.y * sel.y + sel.z * sel.z)));
        if (s > 0.1) {
          ori.push_back(po);
          coef.push_back(P4{s * pa, s * pb, s * pc, s * pd2});
        }
      }
    }
  }
  float srx = sinf(T[0]), crx = cosf(T[0]), sry = sinf(T[1]), cry = cosf(T[1]), srz = sinf(T[2]), crz = cosf(T[2]);
  int n = (int)ori.size();
  ne.n_sel = n;
  ne.A.assign((size_t)n * 6, 0.f);
  ne.B.assign(n, 0.f);
  std::memset(ne.AtA, 0, sizeof(ne.AtA));
  std::memset(ne.AtB, 0, sizeof(ne.AtB));
  if (n < min_rows) return;  // LM:929-932 (min_rows = 50 in the reference; 0 for a shard's partial sums)
  for (int i = 0; i < n; i++) {  // LM:940-964
    const P4& p = ori[i];
    const P4& c = coef[i];
    float arx = (crx * sry * srz * p.x + crx * crz * sry * p.y - srx * sry * p.z) * c.x + (-srx * srz * p.x - crz * srx * p.y - crx * p.z) * c.y +
                (crx * cry * srz * p.x + crx * cry * crz * p.y - cry * srx * p.z) * c.z;
    float ary = ((cry * srx * srz - crz * sry) * p.x + (sry * srz + cry * crz * srx) * p.y + crx * cry * p.z) * c.x +
                ((-cry * crz - srx * sry * srz) * p.x + (cry * srz - crz * srx * sry) * p.y - crx * sry * p.z) * c.z;
    float arz = ((crz * srx * sry - cry * srz) * p.x + (-cry * crz - srx * sry * srz) * p.y) * c.x + (crx * crz * p.x - crx * srz * p.y) * c.y +
                ((sry * srz + cry * crz * srx) * p.x + (crz * sry - cry * srx * srz) * p.y) * c.z;
    float* row = &ne.A[(size_t)i * 6];
    row[0] = arx; row[1] = ary; row[2] = arz; row[3] = c.x; row[4] = c.y; row[5] = c.z;
    ne.B[i] = -c.i;
  }
  normal_equations(ne);
}

struct MapOut {
  bool processed;            // false when the message set was not mapped (never happens with stackFrameNum = 1)
  bool optimised;            // map large enough for the GN loop (LM:749)
  int iterations;
  float transformAftMapped[6];  // /aft_mapped_to_init pose LM:1114-1124
  float transformBefMapped[6];  // smuggled in twist LM:1125-1130
  float transformTobeMapped[6];
  bool surroundPublished;
  Cloud surround;               // LM:1085-1100
  Cloud fullResRegistered;      // LM:1103-1112
  int nCornerStack, nSurfStack, nCornerFromMap, nSurfFromMap;
};

class LaserMapping {
 public:
  static const int W = 21, H = 11, D = 21, NUM = W * H * D;  // LM:72-75
  bool brute = false;
  bool keepClouds = true;  // fill surround / fullResRegistered
  LaserMapping() : cornerArr(NUM), surfArr(NUM) {
    reset_all();
    systemInited = false;
    for (int i = 0; i < 6; i++) Tsum[i] = 0.f;
  }

  void reset_all() {  // LM:434-461 (+ initial values LM:69-71,106-110,418-419)
    gn = GNState();
    for (int i = 0; i < NUM; i++) { cornerArr[i].clear(); surfArr[i].clear(); }
    frameCount = 0;      // stackFrameNum - 1
    mapFrameCount = 4;   // mapFrameNum - 1
    cenW = 10; cenH = 5; cenD = 10;
    for (int i = 0; i < 6; i++) Tincre[i] = Ttobe[i] = Tbef[i] = Taft[i] = 0.f;
  }

  // laserOdometryHandler LM:314-335: called for EVERY odometry message (every sweep), also those that do not lead to a
  // mapping run; drops systemInited when |position| < 1e-6 on all axes.
  void odometry_msg(const float* Tsum_in) {
    if (fabs((double)Tsum_in[3]) < 0.000001 && fabs((double)Tsum_in[4]) < 0.000001 && fabs((double)Tsum_in[5]) < 0.000001) systemInited = false;
    for (int i = 0; i < 6; i++) Tsum[i] = Tsum_in[i];
  }

  // One main-loop body for a synchronised (cornerLast, surfLast, fullRes, odometry) set; uses the latest odometry_msg.
  void process(const Cloud& cornerLast, const Cloud& surfLast, const Cloud& fullRes, MapOut& out) {
    if (!systemInited) {  // LM:434-461
      systemInited = true;
      reset_all();
    }
    out.processed = true;
    out.optimised = false;
    out.iterations = 0;
    out.surroundPublished = false;
    out.surround.clear();
    out.fullResRegistered.clear();

    transform_associate_to_map(Tsum, Tbef, Taft, Tincre, Ttobe);  // LM:465
    Cloud cornerStack2(cornerLast.size()), surfStack2(surfLast.size());
    for (size_t i = 0; i < cornerLast.size(); i++) associate_to_map(Ttobe, cornerLast[i], cornerStack2[i]);
    for (size_t i = 0; i < surfLast.size(); i++) associate_to_map(Ttobe, surfLast[i], surfStack2[i]);

    P4 pointOnYAxis = {0.f, 10.f, 0.f, 0.f};
    associate_to_map(Ttobe, pointOnYAxis, pointOnYAxis);

    // LM:489-495
    int cI = int((Ttobe[3] + 25.0) / 50.0) + cenW;
    int cJ = int((Ttobe[4] + 25.0) / 50.0) + cenH;
    int cK = int((Ttobe[5] + 25.0) / 50.0) + cenD;
    if (Ttobe[3] + 25.0 < 0) cI--;
    if (Ttobe[4] + 25.0 < 0) cJ--;
    if (Ttobe[5] + 25.0 < 0) cK--;

    // LM:497-657: roll the cube grid so the centre stays >= 3 cubes from every face
    while (cI < 3) { shift(0, +1); cI++; cenW++; }
    while (cI >= W - 3) { shift(0, -1); cI--; cenW--; }
    while (cJ < 3) { shift(1, +1); cJ++; cenH++; }
    while (cJ >= H - 3) { shift(1, -1); cJ--; cenH--; }
    while (cK < 3) { shift(2, +1); cK++; cenD++; }
    while (cK >= D - 3) { shift(2, -1); cK--; cenD--; }

    // LM:659-715
    std::vector<int> validInd, surroundInd;
    for (int i = cI - 2; i <= cI + 2; i++)
      for (int j = cJ - 2; j <= cJ + 2; j++)
        for (int k = cK - 2; k <= cK + 2; k++) {
          if (i >= 0 && i < W && j >= 0 && j < H && k >= 0 && k < D) {
            float centerX = (float)(50.0 * (i - cenW));
            float centerY = (float)(50.0 * (j - cenH));
            float centerZ = (float)(50.0 * (k - cenD));
            bool inFOV = false;
            for (int ii = -1; ii <= 1; ii += 2)
              for (int jj = -1; jj <= 1; jj += 2)
                for (int kk = -1; kk <= 1; kk += 2) {
                  float cornerX = (float)(centerX + 25.0 * ii);
                  float cornerY = (float)(centerY + 25.0 * jj);
                  float cornerZ = (float)(centerZ + 25.0 * kk);
                  float s1 = (Ttobe[3] - cornerX) * (Ttobe[3] - cornerX) + (Ttobe[4] - cornerY) * (Ttobe[4] - cornerY) +
                             (Ttobe[5] - cornerZ) * (Ttobe[5] - cornerZ);
                  float s2 = (pointOnYAxis.x - cornerX) * (pointOnYAxis.x - cornerX) + (pointOnYAxis.y - cornerY) * (pointOnYAxis.y - cornerY) +
                             (pointOnYAxis.z - cornerZ) * (pointOnYAxis.z - cornerZ);
                  float check1 = (float)(100.0 + s1 - s2 - 10.0 * sqrt(3.0) * sqrtf(s1));
                  float check2 = (float)(100.0 + s1 - s2 + 10.0 * sqrt(3.0) * sqrtf(s1));
                  if (check1 < 0 && check2 > 0) inFOV = true;
                }
            int ind = i + W * j + W * H * k;
            if (inFOV) validInd.push_back(ind);
            surroundInd.push_back(ind);
          }
        }

    // LM:717-724
    Cloud cornerFromMap, surfFromMap;
    for (int ind : validInd) {
      cornerFromMap.insert(cornerFromMap.end(), cornerArr[ind].begin(), cornerArr[ind].end());
      surfFromMap.insert(surfFromMap.end(), surfArr[ind].begin(), surfArr[ind].end());
    }
    // LM:726-747
    for (auto& p : cornerStack2) associate_tobe_mapped(Ttobe, p, p);
    for (auto& p : surfStack2) associate_tobe_mapped(Ttobe, p, p);
    Cloud cornerStack, surfStack;
    voxel_grid(cornerStack2, 0.2f, cornerStack);
    voxel_grid(surfStack2, 0.4f, surfStack);
    out.nCornerStack = (int)cornerStack.size();
    out.nSurfStack = (int)surfStack.size();
    out.nCornerFromMap = (int)cornerFromMap.size();
    out.nSurfFromMap = (int)surfFromMap.size();

    if (cornerFromMap.size() > 10 && surfFromMap.size() > 100) {  // LM:749
      out.optimised = true;
      KnnIndex kc, ks;
      kc.set(cornerFromMap, brute);
      ks.set(surfFromMap, brute);
      for (int iter = 0; iter < 10; iter++) {
        out.iterations = iter + 1;
        map_iteration(cornerStack, surfStack, cornerFromMap, surfFromMap, kc, ks, Ttobe, nullptr, ne);
        if (ne.n_sel < 50) continue;
        float X[6];
        gn_solve_step(ne.AtA, ne.AtB, iter, 100.f, gn, X);  // LM:968-997
        for (int i = 0; i < 6; i++) Ttobe[i] += X[i];
        float deltaR = (float)sqrt(pow(X[0] * 180.0 / M_PI, 2) + pow(X[1] * 180.0 / M_PI, 2) + pow(X[2] * 180.0 / M_PI, 2));
        float deltaT = (float)sqrt(pow(X[3] * 100, 2) + pow(X[4] * 100, 2) + pow(X[5] * 100, 2));
        if (deltaR < 0.05 && deltaT < 0.05) break;
      }
      for (int i = 0; i < 6; i++) {  // transformUpdate LM:238-241
        Tbef[i] = Tsum[i];
        Taft[i] = Ttobe[i];
      }
    }

    // LM:1023-1059
    insert(cornerStack, cornerArr);
    insert(surfStack, surfArr);
    // LM:1061-1079
    for (int ind : validInd) {
      Cloud ds;
      voxel_grid(cornerArr[ind], 0.2f, ds);
      cornerArr[ind].swap(ds);
      voxel_grid(surfArr[ind], 0.4f, ds);
      surfArr[ind].swap(ds);
    }
    // LM:1081-1101
    mapFrameCount++;
    if (mapFrameCount >= 5) {
      mapFrameCount = 0;
      out.surroundPublished = true;
      if (keepClouds) {
        Cloud s2;
        for (int ind : surroundInd) {
          s2.insert(s2.end(), cornerArr[ind].begin(), cornerArr[ind].end());
          s2.insert(s2.end(), surfArr[ind].begin(), surfArr[ind].end());
        }
        voxel_grid(s2, 0.2f, out.surround);
      }
    }
    // LM:1103-1106
    if (keepClouds) {
      out.fullResRegistered.resize(fullRes.size());
      for (size_t i = 0; i < fullRes.size(); i++) associate_to_map(Ttobe, fullRes[i], out.fullResRegistered[i]);
    }
    for (int i = 0; i < 6; i++) {
      out.transformAftMapped[i] = Taft[i];
      out.transformBefMapped[i] = Tbef[i];
      out.transformTobeMapped[i] = Ttobe[i];
    }
  }

  // state
  std::vector<Cloud> cornerArr, surfArr;
  int frameCount, mapFrameCount;
  int cenW, cenH, cenD;
  float Tsum[6], Tincre[6], Ttobe[6], Tbef[6], Taft[6];
  bool systemInited;
  NormalEq ne;
  GNState gn;

 private:
  // Move every cube one step along `axis` in direction `dir` (+1: contents move towards higher index and the last
  // slab wraps to index 0 emptied; -1: the opposite).  Same effect as the pointer rotations LM:497-657.
  void shift(int axis, int dir) {
    int dims[3] = {W, H, D};
    int n = dims[axis];
    int a1 = (axis + 1) % 3, a2 = (axis + 2) % 3;
    for (int u = 0; u < dims[a1]; u++)
      for (int v = 0; v < dims[a2]; v++) {
        auto idx = [&](int t) {
          int c[3];
          c[axis] = t; c[a1] = u; c[a2] = v;
          return c[0] + W * c[1] + W * H * c[2];
        };
        if (dir > 0) {
          for (int t = n - 1; t >= 1; t--) { cornerArr[idx(t)].swap(cornerArr[idx(t - 1)]); surfArr[idx(t)].swap(surfArr[idx(t - 1)]); }
          cornerArr[idx(0)].clear();
          surfArr[idx(0)].clear();
        } else {
          for (int t = 0; t < n - 1; t++) { cornerArr[idx(t)].swap(cornerArr[idx(t + 1)]); surfArr[idx(t)].swap(surfArr[idx(t + 1)]); }
          cornerArr[idx(n - 1)].clear();
          surfArr[idx(n - 1)].clear();
        }
      }
  }

  void insert(const Cloud& stack, std::vector<Cloud>& arr) {
    P4 sel;
    for (const P4& p : stack) {
      associate_to_map(Ttobe, p, sel);
      int cubeI = int((sel.x + 25.0) / 50.0) + cenW;
      int cubeJ = int((sel.y + 25.0) / 50.0) + cenH;
      int cubeK = int((sel.z + 25.0) / 50.0) + cenD;
      if (sel.x + 25.0 < 0) cubeI--;
      if (sel.y + 25.0 < 0) cubeJ--;
      if (sel.z + 25.0 < 0) cubeK--;
      if (cubeI >= 0 && cubeI < W && cubeJ >= 0 && cubeJ < H && cubeK >= 0 && cubeK < D) arr[cubeI + W * cubeJ + W * H * cubeK].push_back(sel);
    }
  }
};

}  // namespace orc
