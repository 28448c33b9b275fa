// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_linalg.h header).
//
// CPU restatement of the reference's per-sweep feature extraction,
// src/gpsCalibration/src/lidar_slam/loam/scanRegistration.cpp (SR), laserCloudHandler SR:238-752, IMU branch
// excluded (dormant in the shipped pipeline: input_data only replays velodyne_points, so imuPointerLast stays -1).
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

// xyz: n points, `stride` floats apart, sensor frame (x fwd, y left, z up).
inline void extract(const SRParams& prm, SRState& st, const float* xyz, int n, int stride, SROut& out) {
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
