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
  SRM_EMPTY_RING = 8,    // some ring received no point: scanStartInd/EndInd overlap in the reference (SR:480-490)
  SRM_RING_START = 16,   // [n_scans + 1] first index of every ring in the ring-major cloud
  SRM_SCAN_START = 96,   // [n_scans] scanStartInd (SR:484,489)
  SRM_SCAN_END = 160,    // [n_scans] scanEndInd   (SR:485,490)
  SRM_LF_CNT = 224,      // [n_scans] voxel-grid output size per ring
  SRM_PICK_CNT = 288,    // [n_scans][3] sharp / less-sharp / flat picks per ring
  SRM_SIZE = 480
};
enum { SR_PICK_SHARP = 0, SR_PICK_LESS = 96, SR_PICK_FLAT = 216, SR_PICKS_PER_RING = 408 };  // 6*16, 6*20, 6*32

struct SrWs {
  DevBuf ring8, ori_raw, hist, meta;
  DevBuf full, curv, cond, picked, mask_diag, label;
  DevBuf picks, sharp, less_sharp, flat;
  DevBuf lf_valid, lf_tmp, less_flat, segs;
  void release() {
    DevBuf* all[] = {&ring8, &ori_raw, &hist, &meta, &full, &curv, &cond, &picked, &mask_diag, &label,
                     &picks, &sharp, &less_sharp, &flat, &lf_valid, &lf_tmp, &less_flat, &segs};
    for (DevBuf* b : all) b->release();
  }
};

// Enqueues the whole extraction of one sweep on `st`; results and counts (meta) stay on the device.
int lg_extract_launch(SrWs& ws, const SrParams& prm, const float* d_xyz, int n, int stride_bytes, cudaStream_t st, long long* launches);
