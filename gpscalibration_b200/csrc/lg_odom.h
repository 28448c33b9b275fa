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
  DevBuf bounds_c, bounds_s;   // axis-aligned boxes of every 32 consecutive points of the previous sweep's clouds
  bool bounds_valid = false;   // cleared by whoever rewrites those clouds
  void release() {
    DevBuf* all[] = {&best, &c1, &c2, &s1, &s2, &s3, &partials, &ticket, &bounds_c, &bounds_s};
    for (DevBuf* b : all) b->release();
  }
};

int lg_odom_iter_launch(OdomWs& ws, const OdomT& T, const SinCos3& sc, int iter, const float4* sharp, int n_sharp, const float4* flat, int n_flat,
                        const float4* corner_last, int n_cl, const float4* surf_last, int n_sl, double* out28, unsigned long long seq, cudaStream_t st,
                        long long* launches);
// Iterations it0 .. it1-1 of the Gauss-Newton loop (LO:579-1011) WITHOUT the host in between: one thread-block cluster
// computes the rows, reduces them through distributed shared memory, solves the 6x6 system on one thread, updates the
// transform and goes round again.  it0 >= 1 (iteration 0 holds the eigen-decomposition, which stays on the host);
// correspondences are refreshed first when it0 % 5 == 0.  Publishes {T[6], last iteration, converged} to `out`.
struct OdomLoopArgs {
  OdomT T;
  SinCos3 sc;
  float matP[36];
  int degenerate;
  int it0, it1;
};
int lg_odom_loop_launch(OdomWs& ws, const OdomLoopArgs& args, const float4* sharp, int n_sharp, const float4* flat, int n_flat,
                        const float4* corner_last, int n_cl, const float4* surf_last, int n_sl, double* out, unsigned long long seq, cudaStream_t st,
                        long long* launches);
int lg_odom_to_end_launch(const OdomT& T, const SinCos3& sT, const ImuSC& imu, const float4* in0, float4* out0, int n0, const float4* in1,
                          float4* out1, int n1, const float4* in2, float4* out2, int n2, cudaStream_t st, long long* launches);

// ---- batched (lock-step) form: one row of this table per sequence, grid.y of every odometry kernel picks the row ----------
struct OdK {
  const float4 *sharp, *flat, *corner_last, *surf_last;
  int n_sharp, n_flat, n_cl, n_sl;
  float4 *box_c, *sup_c, *box_s, *sup_s;
  unsigned long long* best;
  int *c1, *c2, *s1, *s2, *s3;
  // this round
  int do_bounds, do_refresh, do_iter0, do_loop, do_to_end;
  OdomT T;            // iteration 0 / TransformToEnd
  SinCos3 sc;
  int iter;
  OdomLoopArgs la;    // loop block
  double* out;        // the member's host mailbox (mapped pinned memory)
  unsigned long long seq;
  // TransformToEnd (LO:1087-1106)
  SinCos3 sT;
  ImuSC imu;
  const float4 *in0, *in1, *in2;
  float4 *out0, *out1, *out2;
  int n0, n1, n2;
};
// fills the workspace pointers of `k` (boxes, correspondence arrays) after growing them for these sizes
int lg_odom_batch_prepare(OdomWs& ws, int n_sharp, int n_flat, int n_cl, int n_sl, cudaStream_t st, OdK* k);
bool lg_odom_batch_fits(int n_sharp, int n_flat);  // the cluster kernels hold this many features (else: per-handle path)
int lg_odom_batch_round(const OdK* host_tab, int B, DevBuf& tab, cudaStream_t st, long long* launches);
int lg_odom_batch_to_end(const OdK* host_tab, int B, DevBuf& tab, cudaStream_t st, long long* launches);
