// Host-side interface of lg_extract.cu (scanRegistration kernels).
#pragma once
#include "lg_common.cuh"
#include "lg_voxel.h"

struct SrParams {
  int n_scans;
  int ring_mode;
  float ring_ang_min, ring_ang_step;
  double scan_period;  // SR:56 const double scanPeriod = 0.1
};

// layout of the small int array ("meta") the extraction kernels communicate through
enum {
  SRM_JSTAR = 0,         // first kept input index whose unwrapped azimuth passed pi (halfPassed latch)
  SRM_N_FULL = 1,
  SRM_N_SHARP = 2,
  SRM_N_LESS_SHARP = 3,
  SRM_N_FLAT = 4,
  SRM_N_LESS_FLAT = 5,
  SRM_VOX_OVERFLOW = 6,  // a ring exceeded the shared-memory voxel capacity
  SRM_ERR = 7,           // a sector exceeded the shared-memory sort capacity
  SRM_EMPTY_RING = 8,    // some ring received no point (information only)
  SRM_VIRTUAL = 9,       // some ring's scanStartInd was never written: it spans [0, scanEndInd) and overlaps the rings before
                         // it (SR:480-490) -> the host runs lg_extract_virtual_* after the first pass
  SRM_HEAD = 10,         // ints the host reads back after every sweep
  SRM_RING_START = 16,   // [n_scans + 1] first index of every ring in the ring-major cloud
  SRM_SCAN_START = 96,   // [n_scans] scanStartInd (SR:484,489)
  SRM_SCAN_END = 160,    // [n_scans] scanEndInd   (SR:485,490)
  SRM_LF_CNT = 224,      // [n_scans] voxel-grid output size per ring
  SRM_PICK_CNT = 288,    // [n_scans][3] sharp / less-sharp / flat picks per ring
  SRM_SIZE = 480
};
enum { SR_PICK_SHARP = 0, SR_PICK_LESS = 96, SR_PICK_FLAT = 216, SR_PICKS_PER_RING = 408 };  // 6*16, 6*20, 6*32

// ---- IMU branch (SR:364-434): what the de-skew kernel needs from the node's IMU state (see loam_imu_push) ------------------
constexpr int SR_IMU_Q = 200;  // imuQueLength SR:78
struct SrImuRing {   // the ring of integrated IMU messages, in ring order
  double time[SR_IMU_Q];
  float roll[SR_IMU_Q], pitch[SR_IMU_Q], yaw[SR_IMU_Q];
  float veloX[SR_IMU_Q], veloY[SR_IMU_Q], veloZ[SR_IMU_Q], shiftX[SR_IMU_Q], shiftY[SR_IMU_Q], shiftZ[SR_IMU_Q];
};
struct SrImuCarry {  // persistent between points AND sweeps: Cur, Start, FromStart values, imuPointerFront (SR:76-92)
  float cur[9];      // roll pitch yaw, velo xyz, shift xyz
  float start[9];
  float shift_from_start[3], velo_from_start[3];
  int front;
  int pad;
};
struct SrImuJob {    // one sweep's de-skew: device pointers + scalars
  const SrImuRing* ring;
  SrImuCarry* carry;   // in / out
  int last;            // imuPointerLast
  double time_scan;    // timeScanCur SR:257
};

struct SrWs {
  DevBuf ring8, ori_raw, hist, meta;
  DevBuf full, curv, cond, picked, mask_diag, label;
  DevBuf picks, sharp, less_sharp, flat;
  DevBuf lf_valid, lf_tmp, less_flat, segs;
  DevBuf sort_ind, stale;                          // cloudSortInd; pick flags / labels of the five never re-initialised entries
  DevBuf reach, lf_stage, lf_vout, lf_meta, gkeys;  // virtual-ring pass only
  DevBuf imu_pts, imu_t, imu_fs;                    // IMU branch only: de-skewed points, point times, (front offset, source) per point
  void release() {
    DevBuf* all[] = {&ring8, &ori_raw, &hist, &meta, &full, &curv, &cond, &picked, &mask_diag, &label,
                     &picks, &sharp, &less_sharp, &flat, &lf_valid, &lf_tmp, &less_flat, &segs,
                     &sort_ind, &stale, &reach, &lf_stage, &lf_vout, &lf_meta, &gkeys, &imu_pts, &imu_t, &imu_fs};
    for (DevBuf* b : all) b->release();
  }
};

// Enqueues the whole extraction of one sweep on `st`; results and counts (meta) stay on the device.
// imu != nullptr: the IMU branch runs (the points are de-skewed before they are bucketed; imu->carry is updated on the device).
int lg_extract_launch(SrWs& ws, const SrParams& prm, const float* d_xyz, int n, int stride_bytes, cudaStream_t st, long long* launches,
                      const SrImuJob* imu = nullptr);
// The same for one sweep of each of B sequences with one launch per kernel (grid.y = sequence): members need n > 0, same
// device; `tab` is scratch for the argument table.  Synchronises `st` before it returns.
int lg_extract_launch_batch(SrWs* const* ws, const SrParams* prm, const float* const* d_xyz, const int* n, const int* stride_bytes, int B,
                            DevBuf& tab, cudaStream_t st, long long* launches);
// Sweeps with virtual rings (SRM_VIRTUAL set after the first pass).  lg_extract_virtual_launch replays those rings in the
// reference's serial order; `stage_points` = sum of their scanEndInd (capacity of the less-flat staging).  It leaves
// {offset, count} per ring in ws.lf_meta; the caller voxel-grids ws.lf_stage[offset .. +count) into ws.lf_vout + offset,
// writes the output sizes to meta[SRM_LF_CNT + r] and segs[r].out, and calls lg_extract_finish_launch (feature clouds +
// concatenation, again).
int lg_extract_virtual_launch(SrWs& ws, const SrParams& prm, int n, size_t stage_points, cudaStream_t st, long long* launches);
int lg_extract_finish_launch(SrWs& ws, const SrParams& prm, cudaStream_t st, long long* launches);
