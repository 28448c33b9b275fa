// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_linalg.h header).
//
// CPU restatement of the reference's scan-to-scan odometry,
// src/gpsCalibration/src/lidar_slam/loam/laserOdometry.cpp (LO): TransformToStart LO:123-150, TransformToEnd
// LO:156-227, PluginIMURotation LO:229-287, AccumulateRotation LO:292-309 and the main-loop body LO:502-1147.
// ROS transport is replaced by direct calls; everything else keeps the reference's fp32 expression order.
#pragma once
#include <cmath>
#include <cstring>
#include <vector>

#include "orc_cloud.h"
#include "orc_linalg.h"

namespace orc {

struct ImuTrans {  // SR:730-745 -> LO:385-409; all zero in the shipped pipeline
  float pitchStart = 0, yawStart = 0, rollStart = 0;
  float pitchLast = 0, yawLast = 0, rollLast = 0;
  float shiftX = 0, shiftY = 0, shiftZ = 0;
  float veloX = 0, veloY = 0, veloZ = 0;
};

// LO:123-150
inline void transform_to_start(const float* T, const P4& pi, P4& po) {
  float s = 10 * (pi.i - int(pi.i));
  float rx = s * T[0], ry = s * T[1], rz = s * T[2];
  float tx = s * T[3], ty = s * T[4], tz = s * T[5];
  float x1 = cosf(rz) * (pi.x - tx) + sinf(rz) * (pi.y - ty);
  float y1 = -sinf(rz) * (pi.x - tx) + cosf(rz) * (pi.y - ty);
  float z1 = (pi.z - tz);
  float x2 = x1;
  float y2 = cosf(rx) * y1 + sinf(rx) * z1;
  float z2 = -sinf(rx) * y1 + cosf(rx) * z1;
  po.x = cosf(ry) * x2 - sinf(ry) * z2;
  po.y = y2;
  po.z = sinf(ry) * x2 + cosf(ry) * z2;
  po.i = pi.i;
}

// LO:156-227
inline void transform_to_end(const float* T, const ImuTrans& imu, const P4& pi, P4& po) {
  P4 a;
  transform_to_start(T, pi, a);  // LO:159-180 are the same expressions as LO:126-148
  float x3 = a.x, y3 = a.y, z3 = a.z;
  float rx = T[0], ry = T[1], rz = T[2], tx = T[3], ty = T[4], tz = T[5];
  float x4 = cosf(ry) * x3 + sinf(ry) * z3;
  float y4 = y3;
  float z4 = -sinf(ry) * x3 + cosf(ry) * z3;
  float x5 = x4;
  float y5 = cosf(rx) * y4 - sinf(rx) * z4;
  float z5 = sinf(rx) * y4 + cosf(rx) * z4;
  float x6 = cosf(rz) * x5 - sinf(rz) * y5 + tx;
  float y6 = sinf(rz) * x5 + cosf(rz) * y5 + ty;
  float z6 = z5 + tz;
  float x7 = cosf(imu.rollStart) * (x6 - imu.shiftX) - sinf(imu.rollStart) * (y6 - imu.shiftY);
  float y7 = sinf(imu.rollStart) * (x6 - imu.shiftX) + cosf(imu.rollStart) * (y6 - imu.shiftY);
  float z7 = z6 - imu.shiftZ;
  float x8 = x7;
  float y8 = cosf(imu.pitchStart) * y7 - sinf(imu.pitchStart) * z7;
  float z8 = sinf(imu.pitchStart) * y7 + cosf(imu.pitchStart) * z7;
  float x9 = cosf(imu.yawStart) * x8 + sinf(imu.yawStart) * z8;
  float y9 = y8;
  float z9 = -sinf(imu.yawStart) * x8 + cosf(imu.yawStart) * z8;
  float x10 = cosf(imu.yawLast) * x9 - sinf(imu.yawLast) * z9;
  float y10 = y9;
  float z10 = sinf(imu.yawLast) * x9 + cosf(imu.yawLast) * z9;
  float x11 = x10;
  float y11 = cosf(imu.pitchLast) * y10 + sinf(imu.pitchLast) * z10;
  float z11 = -sinf(imu.pitchLast) * y10 + cosf(imu.pitchLast) * z10;
  po.x = cosf(imu.rollLast) * x11 + sinf(imu.rollLast) * y11;
  po.y = -sinf(imu.rollLast) * x11 + cosf(imu.rollLast) * y11;
  po.z = z11;
  po.i = int(pi.i);
}

// LO:229-287
inline void plugin_imu_rotation(float bcx, float bcy, float bcz, float blx, float bly, float blz, float alx, float aly,
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
inline void accumulate_rotation(float cx, float cy, float cz, float lx, float ly, float lz, float& ox, float& oy, float& oz) {
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

// Point-to-line coefficients shared by LO:688-716 and LM:814-842.  Returns (la, lb, lc, ld2).
inline void line_coeff(float x0, float y0, float z0, float x1, float y1, float z1, float x2, float y2, float z2, float& la,
                       float& lb, float& lc, float& ld2) {
  float cxy = (x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1);
  float cxz = (x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1);
  float cyz = (y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1);
  float a012 = sqrtf(cxy * cxy + cxz * cxz + cyz * cyz);
  float l12 = sqrtf((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2) + (z1 - z2) * (z1 - z2));
  la = ((y1 - y2) * cxy + (z1 - z2) * cxz) / a012 / l12;
  lb = -((x1 - x2) * cxy - (z1 - z2) * cyz) / a012 / l12;
  lc = -((x1 - x2) * cxz + (y1 - y2) * cyz) / a012 / l12;
  ld2 = a012 / l12;
}

struct OdomCorr {  // LO:102-109 (float arrays in the reference; exact integers below 2^24, Appendix B.8)
  std::vector<int> c1, c2, s1, s2, s3;
  void ensure(size_t nc, size_t ns) {
    if (c1.size() < nc) { c1.resize(nc, -1); c2.resize(nc, -1); }
    if (s1.size() < ns) { s1.resize(ns, -1); s2.resize(ns, -1); s3.resize(ns, -1); }
  }
};

struct NormalEq {
  float AtA[36];
  float AtB[6];
  int n_sel;
  std::vector<float> A, B;  // rows (n_sel x 6) and rhs, kept for diagnostics
};

// Build AtA / AtB the way `matAt * matA`, `matAt * matB` do (LO:972-974, LM:965-967).
inline void normal_equations(NormalEq& ne) {
  int n = ne.n_sel;
  for (int i = 0; i < 6; i++) {
    for (int j = 0; j < 6; j++) {
      double s = 0.0;
      for (int t = 0; t < n; t++) s += (double)ne.A[t * 6 + i] * (double)ne.A[t * 6 + j];
      ne.AtA[i * 6 + j] = (float)s;
    }
    double s = 0.0;
    for (int t = 0; t < n; t++) s += (double)ne.A[t * 6 + i] * (double)ne.B[t];
    ne.AtB[i] = (float)s;
  }
}

struct KnnIndex {  // exact kNN over one cloud: brute force is the definition, the kd-tree an asserted-equal accelerator
  const Cloud* cloud = nullptr;
  KdTree tree;
  bool brute = false;
  void set(const Cloud& c, bool use_brute) {
    cloud = &c;
    brute = use_brute;
    if (!brute) tree.build(c);
  }
  int knn(const P4& q, int k, Nbr* out) const { return brute ? knn_brute(*cloud, q, k, out) : tree.knn(q, k, out); }
};

// One pass of the iteration body LO:586-971 (no solve).  `iter` selects refresh (iter % 5 == 0) and weighting (iter >= 5).
inline void odom_iteration(const Cloud& sharp, const Cloud& flat, const Cloud& cornerLast, const Cloud& surfLast,
                           const KnnIndex& kCorner, const KnnIndex& kSurf, const float* T, int iter, OdomCorr& corr,
                           NormalEq& ne) {
  int nSharp = (int)sharp.size(), nFlat = (int)flat.size();
  corr.ensure(nSharp, nFlat);
  std::vector<P4> ori, coef;
  P4 sel;
  for (int i = 0; i < nSharp; i++) {
    transform_to_start(T, sharp[i], sel);
    if (iter % 5 == 0) {
      Nbr nb;
      int found = kCorner.knn(sel, 1, &nb);
      int closest = -1, min2 = -1;
      if (found > 0 && nb.d2 < 25) {
        closest = nb.idx;
        int scan = int(cornerLast[closest].i);
        float d, minD2 = 25;
        // FENCE (i): the reference bounds this scan by cornerPointsSharpNum (LO:620), not by the last cloud's size;
        // clamp so it never reads past laserCloudCornerLast.
        int bound = std::min(nSharp, (int)cornerLast.size());
        for (int j = closest + 1; j < bound; j++) {
          if (int(cornerLast[j].i) > scan + 1.5) break;
          d = (cornerLast[j].x - sel.x) * (cornerLast[j].x - sel.x) + (cornerLast[j].y - sel.y) * (cornerLast[j].y - sel.y) +
              (cornerLast[j].z - sel.z) * (cornerLast[j].z - sel.z);
          if (int(cornerLast[j].i) > scan) {
            if (d < minD2) { minD2 = d; min2 = j; }
          }
        }
        for (int j = closest - 1; j >= 0; j--) {
          if (int(cornerLast[j].i) < scan - 1.5) break;
          d = (cornerLast[j].x - sel.x) * (cornerLast[j].x - sel.x) + (cornerLast[j].y - sel.y) * (cornerLast[j].y - sel.y) +
              (cornerLast[j].z - sel.z) * (cornerLast[j].z - sel.z);
          if (int(cornerLast[j].i) < scan) {
            if (d < minD2) { minD2 = d; min2 = j; }
          }
        }
      }
      corr.c1[i] = closest;
      corr.c2[i] = min2;
    }
    if (corr.c2[i] >= 0) {
      const P4& t1 = cornerLast[corr.c1[i]];
      const P4& t2 = cornerLast[corr.c2[i]];
      float la, lb, lc, ld2;
      line_coeff(sel.x, sel.y, sel.z, t1.x, t1.y, t1.z, t2.x, t2.y, t2.z, la, lb, lc, ld2);
      float s = 1;
      if (iter >= 5) s = (float)(1 - 1.8 * fabsf(ld2));
      if (s > 0.1 && ld2 != 0) {
        ori.push_back(sharp[i]);
        coef.push_back(P4{s * la, s * lb, s * lc, s * ld2});
      }
    }
  }
  for (int i = 0; i < nFlat; i++) {
    transform_to_start(T, flat[i], sel);
    if (iter % 5 == 0) {
      Nbr nb;
      int found = kSurf.knn(sel, 1, &nb);
      int closest = -1, min2 = -1, min3 = -1;
      if (found > 0 && nb.d2 < 25) {
        closest = nb.idx;
        int scan = int(surfLast[closest].i);
        float d, minD2 = 25, minD3 = 25;
        int bound = std::min(nFlat, (int)surfLast.size());  // FENCE (i), LO:776
        for (int j = closest + 1; j < bound; j++) {
          if (int(surfLast[j].i) > scan + 1.5) break;
          d = (surfLast[j].x - sel.x) * (surfLast[j].x - sel.x) + (surfLast[j].y - sel.y) * (surfLast[j].y - sel.y) +
              (surfLast[j].z - sel.z) * (surfLast[j].z - sel.z);
          if (int(surfLast[j].i) <= scan) {
            if (d < minD2) { minD2 = d; min2 = j; }
          } else {
            if (d < minD3) { minD3 = d; min3 = j; }
          }
        }
        for (int j = closest - 1; j >= 0; j--) {
          if (int(surfLast[j].i) < scan - 1.5) break;
          d = (surfLast[j].x - sel.x) * (surfLast[j].x - sel.x) + (surfLast[j].y - sel.y) * (surfLast[j].y - sel.y) +
              (surfLast[j].z - sel.z) * (surfLast[j].z - sel.z);
          if (int(surfLast[j].i) >= scan) {
            if (d < minD2) { minD2 = d; min2 = j; }
          } else {
            if (d < minD3) { minD3 = d; min3 = j; }
          }
        }
      }
      corr.s1[i] = closest;
      corr.s2[i] = min2;
      corr.s3[i] = min3;
    }
    if (corr.s2[i] >= 0 && corr.s3[i] >= 0) {
      const P4& t1 = surfLast[corr.s1[i]];
      const P4& t2 = surfLast[corr.s2[i]];
      const P4& t3 = surfLast[corr.s3[i]];
      float pa = (t2.y - t1.y) * (t3.z - t1.z) - (t3.y - t1.y) * (t2.z - t1.z);
      float pb = (t2.z - t1.z) * (t3.x - t1.x) - (t3.z - t1.z) * (t2.x - t1.x);
      float pc = (t2.x - t1.x) * (t3.y - t1.y) - (t3.x - t1.x) * (t2.y - t1.y);
      float pd = -(pa * t1.x + pb * t1.y + pc * t1.z);
      float ps = sqrtf(pa * pa + pb * pb + pc * pc);
      pa /= ps; pb /= ps; pc /= ps; pd /= ps;
      float pd2 = pa * sel.x + pb * sel.y + pc * sel.z + pd;
      float s = 1;
      if (iter >= 5) s = (float)(1 - 1.8 * fabsf(pd2) / sqrtf(sqrtf(sel.x * sel.x + sel.y * sel.y + sel.z * sel.z)));
      if (s > 0.1 && pd2 != 0) {
        ori.push_back(flat[i]);
        coef.push_back(P4{s * pa, s * pb, s * pc, s * pd2});
      }
    }
  }
  int n = (int)ori.size();
  ne.n_sel = n;
  ne.A.assign((size_t)n * 6, 0.f);
  ne.B.assign(n, 0.f);
  std::memset(ne.AtA, 0, sizeof(ne.AtA));
  std::memset(ne.AtB, 0, sizeof(ne.AtB));
  if (n < 10) return;  // LO:904-907
  // LO:915-971 with s = 1 folded away (multiplying by 1.0f is exact)
  float srx = sinf(T[0]), crx = cosf(T[0]), sry = sinf(T[1]), cry = cosf(T[1]), srz = sinf(T[2]), crz = cosf(T[2]);
  float tx = T[3], ty = T[4], tz = T[5];
  for (int i = 0; i < n; i++) {
    const P4& p = ori[i];
    const P4& c = coef[i];
    float arx = (-crx * sry * srz * p.x + crx * crz * sry * p.y + srx * sry * p.z + tx * crx * sry * srz - ty * crx * crz * sry - tz * srx * sry) * c.x +
                (srx * srz * p.x - crz * srx * p.y + crx * p.z + ty * crz * srx - tz * crx - tx * srx * srz) * c.y +
                (crx * cry * srz * p.x - crx * cry * crz * p.y - cry * srx * p.z + tz * cry * srx + ty * crx * cry * crz - tx * crx * cry * srz) * c.z;
    float ary = ((-crz * sry - cry * srx * srz) * p.x + (cry * crz * srx - sry * srz) * p.y - crx * cry * p.z + tx * (crz * sry + cry * srx * srz) +
                 ty * (sry * srz - cry * crz * srx) + tz * crx * cry) * c.x +
                ((cry * crz - srx * sry * srz) * p.x + (cry * srz + crz * srx * sry) * p.y - crx * sry * p.z + tz * crx * sry -
                 ty * (cry * srz + crz * srx * sry) - tx * (cry * crz - srx * sry * srz)) * c.z;
    float arz = ((-cry * srz - crz * srx * sry) * p.x + (cry * crz - srx * sry * srz) * p.y + tx * (cry * srz + crz * srx * sry) -
                 ty * (cry * crz - srx * sry * srz)) * c.x +
                (-crx * crz * p.x - crx * srz * p.y + ty * crx * srz + tx * crx * crz) * c.y +
                ((cry * crz * srx - sry * srz) * p.x + (crz * sry + cry * srx * srz) * p.y + tx * (sry * srz - cry * crz * srx) -
                 ty * (crz * sry + cry * srx * srz)) * c.z;
    float atx = -(cry * crz - srx * sry * srz) * c.x + crx * srz * c.y - (crz * sry + cry * srx * srz) * c.z;
    float aty = -(cry * srz + crz * srx * sry) * c.x - crx * crz * c.y - (sry * srz - cry * crz * srx) * c.z;
    float atz = crx * sry * c.x - srx * c.y - crx * cry * c.z;
    float* row = &ne.A[(size_t)i * 6];
    row[0] = arx; row[1] = ary; row[2] = arz; row[3] = atx; row[4] = aty; row[5] = atz;
    ne.B[i] = (float)(-0.05 * c.i);
  }
  normal_equations(ne);
}

struct OdomOut {
  float transformSum[6];    // /laser_odom_to_init pose (rx, ry, rz, tx, ty, tz), LO:1059-1064
  float transformation[6];  // sweep-relative transform after the GN loop
  bool odomPublished;       // false only on the (re-)initialisation sweep (LO:519-563 `continue`s before LO:1079)
  bool cloudsPublished;     // LO:541-551 (init) or LO:1126-1146 (every skipFrameNum+1 sweeps)
  bool fullResPublished;    // only LO:1141-1145
  int iterations;           // GN iterations executed
  Cloud cornerLast, surfLast, fullRes;  // what was published (valid when the flags say so)
};

class LaserOdometry {
 public:
  bool brute = false;
  int skipFrameNum = 1;  // LO:52
  LaserOdometry() { reset_all(); }
  void reset_all() {
    systemInited = false;
    frameCount = skipFrameNum;  // LO:495
    for (int i = 0; i < 6; i++) T[i] = Tsum[i] = 0.f;
    cornerLastNum = surfLastNum = 0;
    gn = GNState();
  }
  void control(bool inited) { systemInited = inited; }  // LO:411-415

  // One main-loop body LO:502-1147 for a synchronised message set.
  void process(const Cloud& sharp, const Cloud& lessSharp, const Cloud& flat, const Cloud& lessFlat, const Cloud& fullRes,
               const ImuTrans& imu, OdomOut& out) {
    out.odomPublished = out.cloudsPublished = out.fullResPublished = false;
    out.iterations = 0;
    if (!systemInited) {  // LO:519-563
      cornerLastNum = 0;
      surfLastNum = 0;
      cornerLast = lessSharp;
      surfLast = lessFlat;
      kCorner.set(cornerLast, brute);
      kSurf.set(surfLast, brute);
      out.cornerLast = cornerLast;
      out.surfLast = surfLast;
      out.cloudsPublished = true;
      for (int i = 0; i < 6; i++) T[i] = Tsum[i] = 0.f;
      Tsum[0] += imu.pitchStart;
      Tsum[2] += imu.rollStart;
      systemInited = true;
      for (int i = 0; i < 6; i++) { out.transformSum[i] = Tsum[i]; out.transformation[i] = T[i]; }
      return;
    }
    const float scanPeriod = 0.1f;  // LO:50 (const float)
    T[3] -= imu.veloX * scanPeriod;
    T[4] -= imu.veloY * scanPeriod;
    T[5] -= imu.veloZ * scanPeriod;
    if (cornerLastNum > 10 && surfLastNum > 100) {  // LO:572
      for (int iter = 0; iter < 25; iter++) {
        out.iterations = iter + 1;
        odom_iteration(sharp, flat, cornerLast, surfLast, kCorner, kSurf, T, iter, corr, ne);
        if (ne.n_sel < 10) continue;
        float X[6];
        gn_solve_step(ne.AtA, ne.AtB, iter, 10.f, gn, X);  // LO:975-1004
        for (int i = 0; i < 6; i++) T[i] += X[i];
        for (int i = 0; i < 6; i++)
          if (std::isnan(T[i])) T[i] = 0;
        // LO:1017-1028: rad2deg() and pow() are double (CH `inline double rad2deg(double)`), assigned to float
        float deltaR = (float)sqrt(pow(X[0] * 180.0 / M_PI, 2) + pow(X[1] * 180.0 / M_PI, 2) + pow(X[2] * 180.0 / M_PI, 2));
        float deltaT = (float)sqrt(pow(X[3] * 100, 2) + pow(X[4] * 100, 2) + pow(X[5] * 100, 2));
        if (deltaR < 0.1 && deltaT < 0.1) break;
      }
    }
    // LO:1035-1064
    float rx, ry, rz, tx, ty, tz;
    accumulate_rotation(Tsum[0], Tsum[1], Tsum[2], -T[0], (float)(-T[1] * 1.05), -T[2], rx, ry, rz);
    float x1 = cosf(rz) * (T[3] - imu.shiftX) - sinf(rz) * (T[4] - imu.shiftY);
    float y1 = sinf(rz) * (T[3] - imu.shiftX) + cosf(rz) * (T[4] - imu.shiftY);
    float z1 = (float)(T[5] * 1.05 - imu.shiftZ);
    float x2 = x1;
    float y2 = cosf(rx) * y1 - sinf(rx) * z1;
    float z2 = sinf(rx) * y1 + cosf(rx) * z1;
    tx = Tsum[3] - (cosf(ry) * x2 + sinf(ry) * z2);
    ty = Tsum[4] - y2;
    tz = Tsum[5] - (-sinf(ry) * x2 + cosf(ry) * z2);
    plugin_imu_rotation(rx, ry, rz, imu.pitchStart, imu.yawStart, imu.rollStart, imu.pitchLast, imu.yawLast, imu.rollLast, rx, ry, rz);
    Tsum[0] = rx; Tsum[1] = ry; Tsum[2] = rz; Tsum[3] = tx; Tsum[4] = ty; Tsum[5] = tz;
    out.odomPublished = true;

    // LO:1087-1121
    Cloud newCorner(lessSharp.size()), newSurf(lessFlat.size());
    for (size_t i = 0; i < lessSharp.size(); i++) transform_to_end(T, imu, lessSharp[i], newCorner[i]);
    for (size_t i = 0; i < lessFlat.size(); i++) transform_to_end(T, imu, lessFlat[i], newSurf[i]);
    frameCount++;
    bool pub = frameCount >= skipFrameNum + 1;
    if (pub) {
      out.fullRes.resize(fullRes.size());
      for (size_t i = 0; i < fullRes.size(); i++) transform_to_end(T, imu, fullRes[i], out.fullRes[i]);
    }
    cornerLast.swap(newCorner);
    surfLast.swap(newSurf);
    cornerLastNum = (int)cornerLast.size();
    surfLastNum = (int)surfLast.size();
    if (cornerLastNum > 10 && surfLastNum > 100) {
      kCorner.set(cornerLast, brute);
      kSurf.set(surfLast, brute);
    }
    if (pub) {
      frameCount = 0;
      out.cornerLast = cornerLast;
      out.surfLast = surfLast;
      out.cloudsPublished = true;
      out.fullResPublished = true;
    }
    for (int i = 0; i < 6; i++) { out.transformSum[i] = Tsum[i]; out.transformation[i] = T[i]; }
  }

  // state (the reference's file-scope / main()-scope variables)
  bool systemInited;
  int frameCount;
  float T[6], Tsum[6];
  int cornerLastNum, surfLastNum;
  Cloud cornerLast, surfLast;
  KnnIndex kCorner, kSurf;
  OdomCorr corr;
  NormalEq ne;
  GNState gn;
};

}  // namespace orc
