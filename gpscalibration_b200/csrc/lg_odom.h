// Host-side interface of lg_odom.cu (laserOdometry kernels).
#pragma once
#include <algorithm>

#include "lg_common.cuh"

struct OdomT {  // transformation[6] (LO:111) passed by value to the kernels
  float t[6];
};

struct ImuSC {  // sin/cos of the six /imu_trans angles (host libm) + the start shift, for TransformToEnd LO:201-225
  float s_roll_s, c_roll_s, s_pitch_s, c_pitch_s, s_yaw_s, c_yaw_s;
  float s_roll_l, c_roll_l, s_pitch_l, c_pitch_l, s_yaw_l, c_yaw_l;
  float shift[3];
};

struct OdomWs {
  DevBuf best;                // packed (d2, idx) per query, refreshed every 5th iteration
  DevBuf c1, c2, s1, s2, s3;  // pointSearchCornerInd1/2, pointSearchSurfInd1/2/3 (LO:102-109) as int32
  DevBuf partials, ticket;
  void release() {
    DevBuf* all[] = {&best, &c1, &c2, &s1, &s2, &s3, &partials, &ticket};
    for (DevBuf* b : all) b->release();
  }
};

int lg_odom_iter_launch(OdomWs& ws, const OdomT& T, const SinCos3& sc, int iter, const float4* sharp, int n_sharp, const float4* flat, int n_flat,
                        const float4* corner_last, int n_cl, const float4* surf_last, int n_sl, double* out28, unsigned long long seq, cudaStream_t st,
                        long long* launches);
int lg_odom_to_end_launch(const OdomT& T, const SinCos3& sT, const ImuSC& imu, const float4* in0, float4* out0, int n0, const float4* in1,
                          float4* out1, int n1, const float4* in2, float4* out2, int n2, cudaStream_t st, long long* launches);
