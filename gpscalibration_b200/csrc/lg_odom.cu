// K6 / K8 / K9 — scan-to-scan odometry kernels, the B200 replacement for the hot loops of laserOdometry.cpp:
//   odom_knn_pruned_kernel LO:603,758 exact nearest neighbour of every de-skewed feature in the previous sweep's
//                                    corner / surf cloud (replaces KdTreeFLANN::nearestKSearch(k = 1)): the clouds are
//                                    ring-ordered, so every 32 consecutive points share a tight box; a warp bounds all
//                                    boxes and visits only those that can still hold the answer
//   odom_knn_kernel     (same)       brute force over shared-memory target tiles, (d2, index) packed into one 64-bit
//                                    atomicMin key -- the cross-check (LOAM_ODOM_BRUTE_FORCE=1)
//   odom_corr_kernel    LO:604-677,760-844  the +-1-ring scans for the 2nd / 3rd point, one warp per feature
//   odom_iter_kernel    LO:595,680-971  TransformToStart, point-to-line / point-to-plane coefficients, Jacobian row,
//                                    and the 21 + 6 term reduction (warp shuffle -> CTA -> last-CTA) into a 28-double mailbox
//   odom_to_end_kernel  LO:1087-1106 TransformToEnd over less-sharp, less-flat and (every 2nd sweep) the full cloud
//   odom_loop_kernel    LO:579-1031  iterations 1..4, 5..9, ... of a registration in one launch (solve, convergence test and
//                                    sin/cos of the new angles on the device)
// Iteration 0 (eigen-decomposition, degeneracy projection) and the pose accumulation stay on the host (lg_api.cu).
#include <cooperative_groups.h>

#include "lg_odom.h"
#include "lg_linalg.cuh"
#include "lg_reduce.cuh"

namespace {

// LO:123-150.  sin/cos of the per-point angles s*T[k] are evaluated in fp64 and rounded (lg_sincosf_cr).
__device__ __forceinline__ float4 transform_to_start(const OdomT& T, float4 pi) {
  float s = 10 * (pi.w - int(pi.w));
  float rx = s * T.t[0], ry = s * T.t[1], rz = s * T.t[2];
  float tx = s * T.t[3], ty = s * T.t[4], tz = s * T.t[5];
  float srx, crx, sry, cry, srz, crz;
  lg_sincosf_cr(rx, &srx, &crx);
  lg_sincosf_cr(ry, &sry, &cry);
  lg_sincosf_cr(rz, &srz, &crz);
  float x1 = crz * (pi.x - tx) + srz * (pi.y - ty);
  float y1 = -srz * (pi.x - tx) + crz * (pi.y - ty);
  float z1 = (pi.z - tz);
  float x2 = x1;
  float y2 = crx * y1 + srx * z1;
  float z2 = -srx * y1 + crx * z1;
  float4 po;
  po.x = cry * x2 - sry * z2;
  po.y = y2;
  po.z = sry * x2 + cry * z2;
  po.w = pi.w;
  return po;
}

constexpr int KNN_Q = 128, KNN_T = 512;

__global__ void __launch_bounds__(KNN_Q) odom_knn_kernel(OdomT T, const float4* __restrict__ sharp, int n_sharp, const float4* __restrict__ flat,
                                                          int n_flat, const float4* __restrict__ corner_last, int n_cl,
                                                          const float4* __restrict__ surf_last, int n_sl, int tiles_c,
                                                          unsigned long long* __restrict__ best) {
  __shared__ float4 s_t[KNN_T];
  const bool is_c = (int)blockIdx.x < tiles_c;
  const int tile = is_c ? blockIdx.x : blockIdx.x - tiles_c;
  const float4* q = is_c ? sharp : flat;
  const int nq = is_c ? n_sharp : n_flat;
  const float4* tg = is_c ? corner_last : surf_last;
  const int nt = is_c ? n_cl : n_sl;
  const int t0 = blockIdx.y * KNN_T;
  if (t0 >= nt) return;
  const int cnt = min(KNN_T, nt - t0);
  for (int i = threadIdx.x; i < cnt; i += KNN_Q) s_t[i] = tg[t0 + i];
  __syncthreads();
  const int qi = tile * KNN_Q + threadIdx.x;
  if (qi >= nq) return;
  const float4 sel = transform_to_start(T, q[qi]);
  float bd = __int_as_float(0x7f800000);
  int bi = -1;
#pragma unroll 4
  for (int j = 0; j < cnt; j++) {
    float4 t = s_t[j];
    float d = lg_sqdist(t.x, t.y, t.z, sel.x, sel.y, sel.z);
    if (d < bd) {  // ascending j + strict '<' => smallest index wins ties
      bd = d;
      bi = t0 + j;
    }
  }
  if (bi >= 0) atomicMin(&best[(is_c ? 0 : n_sharp) + qi], lg_pack_nbr(bd, bi));
}

// ---- exact nearest neighbour with box pruning -------------------------------------------------------------------------
// Two levels of axis-aligned boxes over a ring-ordered cloud: a BOX bounds 32 consecutive points, a SUPER-BOX bounds 32
// consecutive boxes (1024 points).  box[2c] / box[2c + 1] = component-wise minimum / maximum (.w of the pair = smallest /
// largest ring id); super[2s], super[2s + 1] likewise.  One CTA of 1024 threads per super-box builds both.
__device__ __forceinline__ void odom_bounds_body(const float4* __restrict__ corner_last, int n_cl, const float4* __restrict__ surf_last,
                                                            int n_sl, float4* __restrict__ box_c, float4* __restrict__ sup_c,
                                                            float4* __restrict__ box_s, float4* __restrict__ sup_s) {
  __shared__ float s_red[8][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int nsup_c = (n_cl + 1023) >> 10;
  const bool is_c = (int)blockIdx.x < nsup_c;
  const int s = is_c ? blockIdx.x : blockIdx.x - nsup_c;
  const float4* pts = is_c ? corner_last : surf_last;
  const int n = is_c ? n_cl : n_sl;
  float4* box = is_c ? box_c : box_s;
  float4* sup = is_c ? sup_c : sup_s;
  const int c = s * 32 + w, j = c * 32 + lane;
  const float inf = __int_as_float(0x7f800000);
  float v[8] = {inf, inf, inf, inf, -inf, -inf, -inf, -inf};  // lo x y z ring, hi x y z ring
  if (j < n) {
    const float4 p = pts[j];
    const float ring = (float)int(p.w);
    v[0] = v[4] = p.x; v[1] = v[5] = p.y; v[2] = v[6] = p.z; v[3] = v[7] = ring;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const float x = __shfl_xor_sync(0xffffffffu, v[k], o);
      v[k] = k < 4 ? fminf(v[k], x) : fmaxf(v[k], x);
    }
  if (lane == 0) {
    if (c * 32 < n) {
      box[2 * c] = make_float4(v[0], v[1], v[2], v[3]);
      box[2 * c + 1] = make_float4(v[4], v[5], v[6], v[7]);
    }
#pragma unroll
    for (int k = 0; k < 8; k++) s_red[k][w] = v[k];
  }
  __syncthreads();
  if (w == 0) {
    float r[8];
#pragma unroll
    for (int k = 0; k < 8; k++) r[k] = s_red[k][lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const float x = __shfl_xor_sync(0xffffffffu, r[k], o);
        r[k] = k < 4 ? fminf(r[k], x) : fmaxf(r[k], x);
      }
    if (lane == 0) {
      sup[2 * s] = make_float4(r[0], r[1], r[2], r[3]);
      sup[2 * s + 1] = make_float4(r[4], r[5], r[6], r[7]);
    }
  }
}

// squared distance from s to the box [lo, hi] (0 inside)
__device__ __forceinline__ float box_d2(float4 s, float4 lo, float4 hi) {
  const float dx = fmaxf(fmaxf(lo.x - s.x, s.x - hi.x), 0.f);
  const float dy = fmaxf(fmaxf(lo.y - s.y, s.y - hi.y), 0.f);
  const float dz = fmaxf(fmaxf(lo.z - s.z, s.z - hi.z), 0.f);
  return dx * dx + dy * dy + dz * dz;
}

// One WARP per feature.  Lower-bound the super-boxes, descend into the most promising one first, then into every other
// super-box / box whose bound (shrunk by 1e-5 relative, far more than the rounding of either distance) does not exceed the
// best distance so far.  Everything skipped holds only points that are strictly farther, so the result -- the minimum of
// the same packed (fp32 d2, index) key as the brute-force kernel -- is identical, ties included.
constexpr int KP_WARPS = 8;
__device__ __forceinline__ unsigned long long odom_knn_pruned_warp(const OdomT& T, int q, int lane, const float4* __restrict__ sharp, int n_sharp,
                                                                   const float4* __restrict__ flat, const float4* __restrict__ corner_last, int n_cl,
                                                                   const float4* __restrict__ surf_last, int n_sl,
                                                                   const float4* __restrict__ box_c, const float4* __restrict__ sup_c,
                                                                   const float4* __restrict__ box_s, const float4* __restrict__ sup_s) {
  const bool is_c = q < n_sharp;
  const float4* pts = is_c ? corner_last : surf_last;
  const float4* box = is_c ? box_c : box_s;
  const float4* sup = is_c ? sup_c : sup_s;
  const int nt = is_c ? n_cl : n_sl;
  const int nbox = (nt + 31) >> 5, nsup = (nt + 1023) >> 10;
  const float4 sel = transform_to_start(T, is_c ? sharp[q] : flat[q - n_sharp]);
  const float inf = __int_as_float(0x7f800000);
  unsigned long long bestkey = ~0ull;
  auto warp_min_key = [&](unsigned long long k) {
    const unsigned int hi = (unsigned int)(k >> 32);
    const unsigned int mhi = __reduce_min_sync(0xffffffffu, hi);
    const unsigned int lo = (hi == mhi) ? (unsigned int)k : 0xffffffffu;
    const unsigned int mlo = __reduce_min_sync(0xffffffffu, lo);
    return ((unsigned long long)mhi << 32) | mlo;
  };
  auto open_box = [&](int c) {
    const int j = c * 32 + lane;
    unsigned long long k = ~0ull;
    if (j < nt) {
      const float4 t = pts[j];
      k = lg_pack_nbr(lg_sqdist(t.x, t.y, t.z, sel.x, sel.y, sel.z), j);
    }
    bestkey = min(bestkey, warp_min_key(k));
  };
  // all boxes of super-box s that can still matter, the closest-looking one first
  auto open_super = [&](int s) {
    const int c = s * 32 + lane;
    float lb = inf;
    if (c < nbox) lb = box_d2(sel, box[2 * c], box[2 * c + 1]) * 0.99999f;
    const unsigned int mlb = __reduce_min_sync(0xffffffffu, __float_as_uint(lb));  // non-negative floats order like their bits
    unsigned int m = __ballot_sync(0xffffffffu, lb <= lg_nbr_d2(bestkey) || bestkey == ~0ull);
    const unsigned int mfirst = __ballot_sync(0xffffffffu, __float_as_uint(lb) == mlb);
    if (mfirst & m) {
      const int p = __ffs(mfirst & m) - 1;
      m &= ~(1u << p);
      open_box(s * 32 + p);
    }
    while (m) {
      const int p = __ffs(m) - 1;
      m &= m - 1;
      if (__shfl_sync(0xffffffffu, lb, p) <= lg_nbr_d2(bestkey)) open_box(s * 32 + p);  // the best may have improved meanwhile
    }
  };
  // the most promising super-box first
  unsigned long long mine = ~0ull;
  for (int s = lane; s < nsup; s += 32) {
    const float lb = box_d2(sel, sup[2 * s], sup[2 * s + 1]);
    mine = min(mine, ((unsigned long long)__float_as_uint(lb) << 32) | (unsigned int)s);
  }
  const int first = nsup > 0 ? (int)(unsigned int)warp_min_key(mine) : -1;
  if (first >= 0) open_super(first);
  for (int base = 0; base < nsup; base += 32) {
    const int s = base + lane;
    float lb = inf;
    if (s < nsup && s != first) lb = box_d2(sel, sup[2 * s], sup[2 * s + 1]) * 0.99999f;
    unsigned int m = __ballot_sync(0xffffffffu, lb <= lg_nbr_d2(bestkey));
    while (m) {
      const int p = __ffs(m) - 1;
      m &= m - 1;
      if (__shfl_sync(0xffffffffu, lb, p) <= lg_nbr_d2(bestkey)) open_super(base + p);
    }
  }
  return bestkey;
}

__device__ __forceinline__ float sqd(float4 a, float4 sel) {
  return (a.x - sel.x) * (a.x - sel.x) + (a.y - sel.y) * (a.y - sel.y) + (a.z - sel.z) * (a.z - sel.z);
}

// Point-to-line coefficients shared with mapping (LO:688-716 == LM:814-842).
__device__ __forceinline__ void line_coeff(float x0, float y0, float z0, float x1, float y1, float z1, float x2, float y2, float z2,
                                           float& la, float& lb, float& lc, float& ld2) {
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


__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned int lo = __shfl_xor_sync(0xffffffffu, (unsigned int)v, o);
    unsigned int hi = __shfl_xor_sync(0xffffffffu, (unsigned int)(v >> 32), o);
    unsigned long long t = ((unsigned long long)hi << 32) | lo;
    v = t < v ? t : v;
  }
  return v;
}

constexpr int CORR_WARPS = 8;
constexpr unsigned long long NONE64 = ~0ull;

// LO:604-677 / LO:760-844, one WARP per feature: the sequential +-1-ring scans of the reference become 32-wide
// chunks.  The reference's `break` (first j whose ring leaves the window) is reproduced literally: lanes before the
// first breaking lane of a chunk are candidates, the chunk containing it is the last one.  Tie rule of the
// reference's strict '<' updates: forward scan first (smallest j wins ties), backward scan only replaces on a strictly
// smaller distance (largest j wins ties among backward candidates).
__global__ void __launch_bounds__(CORR_WARPS * 32) odom_corr_kernel(OdomT T, const float4* __restrict__ sharp, int n_sharp,
                                                                     const float4* __restrict__ flat, int n_flat,
                                                                     const float4* __restrict__ corner_last, int n_cl,
                                                                     const float4* __restrict__ surf_last, int n_sl,
                                                                     const unsigned long long* __restrict__ best, int* __restrict__ c1,
                                                                     int* __restrict__ c2, int* __restrict__ s1, int* __restrict__ s2,
                                                                     int* __restrict__ s3) {
  const int lane = threadIdx.x & 31;
  const int q = blockIdx.x * CORR_WARPS + (threadIdx.x >> 5);
  if (q >= n_sharp + n_flat) return;
  const bool is_c = q < n_sharp;
  const int f = is_c ? q : q - n_sharp;
  const float4* pts = is_c ? corner_last : surf_last;
  const int nlast = is_c ? n_cl : n_sl;
  const int bound = min(is_c ? n_sharp : n_flat, nlast);  // FENCE (i): LO:620 / LO:776 bound by the CURRENT feature count
  const float4 sel = transform_to_start(T, is_c ? sharp[f] : flat[f]);
  const unsigned long long b = best[q];
  int closest = -1, r2 = -1, r3 = -1;
  if (b != NONE64 && lg_nbr_d2(b) < 25) {
    closest = lg_nbr_idx(b);
    const int scan = int(pts[closest].w);
    unsigned long long f2 = NONE64, f3 = NONE64, b2 = NONE64, b3 = NONE64;
    // four 32-wide chunks per step: the loads of a step are independent, so four L2 round trips overlap
    for (int base = closest + 1; base < bound; base += 128) {
      float4 t[4];
      unsigned int bm[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int j = base + 32 * u + lane;
        t[u] = (j < bound) ? pts[j] : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      bool stop = false;
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int j = base + 32 * u + lane;
        const bool valid = j < bound;
        const int r = int(t[u].w);
        bm[u] = __ballot_sync(0xffffffffu, valid && (r > scan + 1.5));
        const bool live = !stop && valid && (bm[u] == 0u || lane < (__ffs(bm[u]) - 1));
        if (live) {
          float d = sqd(t[u], sel);
          if (d < 25) {
            unsigned long long key = lg_pack_nbr(d, j);
            if (is_c) {
              if (r > scan) f2 = min(f2, key);
            } else {
              if (r <= scan) f2 = min(f2, key); else f3 = min(f3, key);
            }
          }
        }
        if (bm[u] != 0u) stop = true;
      }
      if (stop) break;
    }
    for (int base = closest - 1; base >= 0; base -= 128) {
      float4 t[4];
      unsigned int bm[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int j = base - 32 * u - lane;
        t[u] = (j >= 0) ? pts[j] : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      bool stop = false;
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int j = base - 32 * u - lane;
        const bool valid = j >= 0;
        const int r = int(t[u].w);
        bm[u] = __ballot_sync(0xffffffffu, valid && (r < scan - 1.5));
        const bool live = !stop && valid && (bm[u] == 0u || lane < (__ffs(bm[u]) - 1));
        if (live) {
          float d = sqd(t[u], sel);
          if (d < 25) {
            unsigned long long key = lg_pack_nbr(d, 0x7fffffff - j);  // first met (largest j) wins ties
            if (is_c) {
              if (r < scan) b2 = min(b2, key);
            } else {
              if (r >= scan) b2 = min(b2, key); else b3 = min(b3, key);
            }
          }
        }
        if (bm[u] != 0u) stop = true;
      }
      if (stop) break;
    }
    f2 = warp_min_u64(f2); b2 = warp_min_u64(b2);
    if (f2 != NONE64 && (b2 == NONE64 || !(lg_nbr_d2(b2) < lg_nbr_d2(f2)))) r2 = lg_nbr_idx(f2);
    else if (b2 != NONE64) r2 = 0x7fffffff - lg_nbr_idx(b2);
    if (!is_c) {
      f3 = warp_min_u64(f3); b3 = warp_min_u64(b3);
      if (f3 != NONE64 && (b3 == NONE64 || !(lg_nbr_d2(b3) < lg_nbr_d2(f3)))) r3 = lg_nbr_idx(f3);
      else if (b3 != NONE64) r3 = 0x7fffffff - lg_nbr_idx(b3);
    }
  }
  if (lane == 0) {
    if (is_c) {
      c1[f] = closest;
      c2[f] = r2;
    } else {
      s1[f] = closest;
      s2[f] = r2;
      s3[f] = r3;
    }
  }
}

// The same ring scans with box pruning.  The scan windows are index ranges of the ring-ordered cloud, so they are walked
// box by box (32 boxes bounded per step, one per lane): a box is opened only if it may hold a candidate that beats the
// current best of its category (or is the one in which the reference's `break` fires, recognised by its ring range);
// everything else holds only points at 5 m or more, or strictly farther than the best so far.  Inside an opened box the
// literal rules of odom_corr_kernel apply, so the result is identical (tests compare both against the oracle).
__device__ __forceinline__ void odom_corr_pruned_warp(const OdomT& T, int q, int lane, unsigned long long b, const float4* __restrict__ sharp,
                                                      int n_sharp, const float4* __restrict__ flat, int n_flat,
                                                      const float4* __restrict__ corner_last, int n_cl, const float4* __restrict__ surf_last,
                                                      int n_sl, const float4* __restrict__ box_c, const float4* __restrict__ box_s,
                                                      int* __restrict__ c1, int* __restrict__ c2, int* __restrict__ s1, int* __restrict__ s2,
                                                      int* __restrict__ s3) {
  const bool is_c = q < n_sharp;
  const int f = is_c ? q : q - n_sharp;
  const float4* pts = is_c ? corner_last : surf_last;
  const float4* box = is_c ? box_c : box_s;
  const int nlast = is_c ? n_cl : n_sl;
  const int bound = min(is_c ? n_sharp : n_flat, nlast);  // FENCE (i): LO:620 / LO:776 bound by the CURRENT feature count
  const float4 sel = transform_to_start(T, is_c ? sharp[f] : flat[f]);
  const float inf = __int_as_float(0x7f800000);
  int closest = -1, r2 = -1, r3 = -1;
  if (b != NONE64 && lg_nbr_d2(b) < 25) {
    closest = lg_nbr_idx(b);
    const int scan = int(pts[closest].w);
    const float fscan = (float)scan;
    unsigned long long f2 = NONE64, f3 = NONE64, b2 = NONE64, b3 = NONE64;
    auto dist_of = [&](unsigned long long k) {  // warp-wide best distance of a category (inf while it is empty)
      const unsigned int hi = (unsigned int)(k >> 32);
      const unsigned int mhi = __reduce_min_sync(0xffffffffu, hi);
      return mhi == 0xffffffffu ? inf : __uint_as_float(mhi);
    };
    // ---- forward: j = closest + 1 .. , stops at the first j whose ring exceeds scan + 1.5 or at `bound`
    {
      float F2 = inf, F3 = inf;
      bool stop = false;
      for (int cbase = (closest + 1) >> 5; !stop && cbase * 32 < bound; cbase += 32) {
        const int cb = cbase + lane;
        const bool inr = cb * 32 < bound;
        float4 lo = make_float4(0.f, 0.f, 0.f, 0.f), hi = lo;
        if (inr) {
          lo = box[2 * cb];
          hi = box[2 * cb + 1];
        }
        const unsigned int brkm = __ballot_sync(0xffffffffu, inr && hi.w > fscan + 1.5f);
        const int last = brkm ? __ffs(brkm) - 1 : 31;  // boxes behind the one in which the scan breaks are never reached
        const float lbm = inr ? box_d2(sel, lo, hi) * 0.99999f : inf;
        bool need = lbm < 25.f;
        if (is_c) need = need && hi.w > fscan && lbm <= F2;
        else need = need && ((lo.w <= fscan && lbm <= F2) || (hi.w > fscan && lbm <= F3));
        // the box in which the scan may break is always opened: only the points themselves tell whether it does
        need = inr && lane <= last && (need || ((brkm >> lane) & 1u));
        unsigned int m = __ballot_sync(0xffffffffu, need);
        bool broke = false;
        while (m) {
          const int p = __ffs(m) - 1;
          m &= m - 1;
          const bool flagged = (brkm >> p) & 1u;
          if (!flagged && __shfl_sync(0xffffffffu, lbm, p) > fmaxf(F2, is_c ? F2 : F3)) continue;  // pruned by an improvement meanwhile
          const int j = (cbase + p) * 32 + lane;
          const bool valid = j > closest && j < bound;
          const float4 t = valid ? pts[j] : make_float4(0.f, 0.f, 0.f, 0.f);
          const int r = int(t.w);
          const unsigned int bm = __ballot_sync(0xffffffffu, valid && (r > scan + 1.5));
          const bool live = valid && (bm == 0u || lane < (__ffs(bm) - 1));
          if (live) {
            const float d = sqd(t, sel);
            if (d < 25) {
              const unsigned long long key = lg_pack_nbr(d, j);
              if (is_c) {
                if (r > scan) f2 = min(f2, key);
              } else {
                if (r <= scan) f2 = min(f2, key); else f3 = min(f3, key);
              }
            }
          }
          F2 = dist_of(f2);
          if (!is_c) F3 = dist_of(f3);
          if (bm != 0u) broke = true;
        }
        if (broke) stop = true;
        else if (brkm) cbase += last + 1 - 32;  // the flagged box did not break after all (rings out of order): go on behind it
      }
    }
    // ---- backward: j = closest - 1 .. 0, stops at the first j whose ring falls below scan - 1.5
    {
      float B2 = inf, B3 = inf;
      bool stop = false;
      for (int ctop = (closest - 1) >> 5; !stop && ctop >= 0 && closest >= 1; ctop -= 32) {
        const int cb = ctop - lane;
        const bool inr = cb >= 0;
        float4 lo = make_float4(0.f, 0.f, 0.f, 0.f), hi = lo;
        if (inr) {
          lo = box[2 * cb];
          hi = box[2 * cb + 1];
        }
        const unsigned int brkm = __ballot_sync(0xffffffffu, inr && lo.w < fscan - 1.5f);
        const int last = brkm ? __ffs(brkm) - 1 : 31;
        const float lbm = inr ? box_d2(sel, lo, hi) * 0.99999f : inf;
        bool need = lbm < 25.f;
        if (is_c) need = need && lo.w < fscan && lbm <= B2;
        else need = need && ((hi.w >= fscan && lbm <= B2) || (lo.w < fscan && lbm <= B3));
        need = inr && lane <= last && (need || ((brkm >> lane) & 1u));
        unsigned int m = __ballot_sync(0xffffffffu, need);
        bool broke = false;
        while (m) {
          const int p = __ffs(m) - 1;
          m &= m - 1;
          const bool flagged = (brkm >> p) & 1u;
          if (!flagged && __shfl_sync(0xffffffffu, lbm, p) > fmaxf(B2, is_c ? B2 : B3)) continue;
          const int j = (ctop - p) * 32 + lane;
          const bool valid = j < closest && j >= 0;
          const float4 t = valid ? pts[j] : make_float4(0.f, 0.f, 0.f, 0.f);
          const int r = int(t.w);
          const unsigned int bm = __ballot_sync(0xffffffffu, valid && (r < scan - 1.5));
          const bool live = valid && (bm == 0u || lane > (31 - __clz(bm)));  // descending j: the highest violating lane breaks
          if (live) {
            const float d = sqd(t, sel);
            if (d < 25) {
              const unsigned long long key = lg_pack_nbr(d, 0x7fffffff - j);  // first met (largest j) wins ties
              if (is_c) {
                if (r < scan) b2 = min(b2, key);
              } else {
                if (r >= scan) b2 = min(b2, key); else b3 = min(b3, key);
              }
            }
          }
          B2 = dist_of(b2);
          if (!is_c) B3 = dist_of(b3);
          if (bm != 0u) broke = true;
        }
        if (broke) stop = true;
        else if (brkm) ctop -= last + 1 - 32;
      }
    }
    f2 = warp_min_u64(f2); b2 = warp_min_u64(b2);
    if (f2 != NONE64 && (b2 == NONE64 || !(lg_nbr_d2(b2) < lg_nbr_d2(f2)))) r2 = lg_nbr_idx(f2);
    else if (b2 != NONE64) r2 = 0x7fffffff - lg_nbr_idx(b2);
    if (!is_c) {
      f3 = warp_min_u64(f3); b3 = warp_min_u64(b3);
      if (f3 != NONE64 && (b3 == NONE64 || !(lg_nbr_d2(b3) < lg_nbr_d2(f3)))) r3 = lg_nbr_idx(f3);
      else if (b3 != NONE64) r3 = 0x7fffffff - lg_nbr_idx(b3);
    }
  }
  if (lane == 0) {
    if (is_c) {
      c1[f] = closest;
      c2[f] = r2;
    } else {
      s1[f] = closest;
      s2[f] = r2;
      s3[f] = r3;
    }
  }
}

// One refresh of the correspondences (LO:598-677, 756-844) in one launch, one WARP per feature: exact nearest neighbour
// over the box hierarchy, then the +-1-ring scans around it -- the key never leaves the warp's registers (it is still
// written to `best` for the brute-force cross-check and diagnostics).
__global__ void __launch_bounds__(KP_WARPS * 32) odom_refresh_pruned_kernel(OdomT T, const float4* __restrict__ sharp, int n_sharp,
                                                                             const float4* __restrict__ flat, int n_flat,
                                                                             const float4* __restrict__ corner_last, int n_cl,
                                                                             const float4* __restrict__ surf_last, int n_sl,
                                                                             const float4* __restrict__ box_c, const float4* __restrict__ sup_c,
                                                                             const float4* __restrict__ box_s, const float4* __restrict__ sup_s,
                                                                             unsigned long long* __restrict__ best, int* __restrict__ c1,
                                                                             int* __restrict__ c2, int* __restrict__ s1, int* __restrict__ s2,
                                                                             int* __restrict__ s3) {
  const int lane = threadIdx.x & 31;
  const int q = blockIdx.x * KP_WARPS + (threadIdx.x >> 5);
  if (q >= n_sharp + n_flat) return;
  const unsigned long long b = odom_knn_pruned_warp(T, q, lane, sharp, n_sharp, flat, corner_last, n_cl, surf_last, n_sl, box_c, sup_c, box_s, sup_s);
  if (lane == 0) best[q] = b;
  odom_corr_pruned_warp(T, q, lane, b, sharp, n_sharp, flat, n_flat, corner_last, n_cl, surf_last, n_sl, box_c, box_s, c1, c2, s1, s2, s3);
}

constexpr int IT_NT = 128;

// One feature's contribution to the normal equations, in two parts.  odom_row_load: everything that does not depend on
// the transform -- the feature point, its correspondence points (line, LO:680-687) or the normalised plane through them
// (LO:847-866); odom_row_eval: TransformToStart, point-to-line / point-to-plane distance and weight (LO:688-746 /
// LO:867-901), Jacobian row (LO:915-971), accumulated into `acc`.  A kernel that iterates (odom_loop_kernel) loads once.
struct OdomRowIn {
  float4 ori;   // the feature point
  float4 a, b;  // line: the two correspondence points; plane: a = {pa, pb, pc, pd}
  int kind;     // 0 = no correspondence, 1 = line, 2 = plane
};
__device__ __forceinline__ OdomRowIn odom_row_load(int q, const float4* __restrict__ sharp, int n_sharp, const float4* __restrict__ flat, int n_flat,
                                                   const float4* __restrict__ corner_last, const float4* __restrict__ surf_last,
                                                   const int* __restrict__ c1, const int* __restrict__ c2, const int* __restrict__ s1,
                                                   const int* __restrict__ s2, const int* __restrict__ s3) {
  OdomRowIn in;
  in.kind = 0;
  in.ori = in.a = in.b = make_float4(0.f, 0.f, 0.f, 0.f);
  if (q < n_sharp) {
    in.ori = sharp[q];
    const int i2 = c2[q];
    if (i2 >= 0) {
      in.a = corner_last[c1[q]];
      in.b = corner_last[i2];
      in.kind = 1;
    }
  } else if (q < n_sharp + n_flat) {
    const int f = q - n_sharp;
    in.ori = flat[f];
    const int i2 = s2[f], i3 = s3[f];
    if (i2 >= 0 && i3 >= 0) {  // LO:847-866
      float4 t1 = surf_last[s1[f]], t2 = surf_last[i2], t3 = surf_last[i3];
      float pa = (t2.y - t1.y) * (t3.z - t1.z) - (t3.y - t1.y) * (t2.z - t1.z);
      float pb = (t2.z - t1.z) * (t3.x - t1.x) - (t3.z - t1.z) * (t2.x - t1.x);
      float pc = (t2.x - t1.x) * (t3.y - t1.y) - (t3.x - t1.x) * (t2.y - t1.y);
      float pd = -(pa * t1.x + pb * t1.y + pc * t1.z);
      float ps = sqrtf(pa * pa + pb * pb + pc * pc);
      pa /= ps; pb /= ps; pc /= ps; pd /= ps;
      in.a = make_float4(pa, pb, pc, pd);
      in.kind = 2;
    }
  }
  return in;
}

__device__ __forceinline__ void odom_row_eval(const OdomRowIn& in, const OdomT& T, const SinCos3& sc, int iter, Acc28& acc) {
  if (in.kind == 0) return;
  const float4 ori = in.ori;
  const float4 sel = transform_to_start(T, ori);
  float4 coef;
  bool keep;
  if (in.kind == 1) {  // LO:688-746
    float la, lb, lc, ld2;
    line_coeff(sel.x, sel.y, sel.z, in.a.x, in.a.y, in.a.z, in.b.x, in.b.y, in.b.z, la, lb, lc, ld2);
    float s = 1;
    if (iter >= 5) s = (float)(1 - 1.8 * fabsf(ld2));
    coef = make_float4(s * la, s * lb, s * lc, s * ld2);
    keep = (s > 0.1 && ld2 != 0);
  } else {  // LO:867-901
    const float pa = in.a.x, pb = in.a.y, pc = in.a.z, pd = in.a.w;
    float pd2 = pa * sel.x + pb * sel.y + pc * sel.z + pd;
    float s = 1;
    if (iter >= 5) s = (float)(1 - 1.8 * fabsf(pd2) / sqrtf(sqrtf(sel.x * sel.x + sel.y * sel.y + sel.z * sel.z)));
    coef = make_float4(s * pa, s * pb, s * pc, s * pd2);
    keep = (s > 0.1 && pd2 != 0);
  }
  if (keep) {  // LO:915-971 with s = 1 folded away (x * 1.0f is exact)
    const float srx = sc.srx, crx = sc.crx, sry = sc.sry, cry = sc.cry, srz = sc.srz, crz = sc.crz;
    const float tx = T.t[3], ty = T.t[4], tz = T.t[5];
    const float4 p = ori, c = coef;
    float a[6];
    a[0] = (-crx * sry * srz * p.x + crx * crz * sry * p.y + srx * sry * p.z + tx * crx * sry * srz - ty * crx * crz * sry - tz * srx * sry) * c.x +
           (srx * srz * p.x - crz * srx * p.y + crx * p.z + ty * crz * srx - tz * crx - tx * srx * srz) * c.y +
           (crx * cry * srz * p.x - crx * cry * crz * p.y - cry * srx * p.z + tz * cry * srx + ty * crx * cry * crz - tx * crx * cry * srz) * c.z;
    a[1] = ((-crz * sry - cry * srx * srz) * p.x + (cry * crz * srx - sry * srz) * p.y - crx * cry * p.z + tx * (crz * sry + cry * srx * srz) +
            ty * (sry * srz - cry * crz * srx) + tz * crx * cry) * c.x +
           ((cry * crz - srx * sry * srz) * p.x + (cry * srz + crz * srx * sry) * p.y - crx * sry * p.z + tz * crx * sry -
            ty * (cry * srz + crz * srx * sry) - tx * (cry * crz - srx * sry * srz)) * c.z;
    a[2] = ((-cry * srz - crz * srx * sry) * p.x + (cry * crz - srx * sry * srz) * p.y + tx * (cry * srz + crz * srx * sry) -
            ty * (cry * crz - srx * sry * srz)) * c.x +
           (-crx * crz * p.x - crx * srz * p.y + ty * crx * srz + tx * crx * crz) * c.y +
           ((cry * crz * srx - sry * srz) * p.x + (crz * sry + cry * srx * srz) * p.y + tx * (sry * srz - cry * crz * srx) -
            ty * (crz * sry + cry * srx * srz)) * c.z;
    a[3] = -(cry * crz - srx * sry * srz) * c.x + crx * srz * c.y - (crz * sry + cry * srx * srz) * c.z;
    a[4] = -(cry * srz + crz * srx * sry) * c.x - crx * crz * c.y - (sry * srz - cry * crz * srx) * c.z;
    a[5] = crx * sry * c.x - srx * c.y - crx * cry * c.z;
    float b = (float)(-0.05 * c.w);
    acc.add_row(a, b);
  }
}

__device__ __forceinline__ void odom_row(int q, const OdomT& T, const SinCos3& sc, int iter, const float4* __restrict__ sharp, int n_sharp,
                                         const float4* __restrict__ flat, int n_flat, const float4* __restrict__ corner_last,
                                         const float4* __restrict__ surf_last, const int* __restrict__ c1, const int* __restrict__ c2,
                                         const int* __restrict__ s1, const int* __restrict__ s2, const int* __restrict__ s3, Acc28& acc) {
  const OdomRowIn in = odom_row_load(q, sharp, n_sharp, flat, n_flat, corner_last, surf_last, c1, c2, s1, s2, s3);
  odom_row_eval(in, T, sc, iter, acc);
}

__global__ void __launch_bounds__(IT_NT) odom_iter_kernel(OdomT T, SinCos3 sc, int iter, const float4* __restrict__ sharp, int n_sharp,
                                                           const float4* __restrict__ flat, int n_flat, const float4* __restrict__ corner_last,
                                                           int n_cl, const float4* __restrict__ surf_last, int n_sl,
                                                           const int* __restrict__ c1, const int* __restrict__ c2,
                                                           const int* __restrict__ s1, const int* __restrict__ s2, const int* __restrict__ s3,
                                                           double* __restrict__ partials, unsigned int* __restrict__ ticket, double* __restrict__ out28,
                                                           unsigned long long seq) {
  Acc28 acc;
  acc.clear();
  const int q = blockIdx.x * IT_NT + threadIdx.x;
  odom_row(q, T, sc, iter, sharp, n_sharp, flat, n_flat, corner_last, surf_last, c1, c2, s1, s2, s3, acc);
  lg_reduce28<IT_NT>(acc, partials, ticket, out28, seq);
}

// Same iteration as ONE thread-block cluster (8 CTAs x 256 threads): every CTA reduces its rows to 28 doubles in shared
// memory, the cluster synchronises once and CTA 0 adds the eight partials straight out of its peers' shared memory
// (DSMEM) in rank order before publishing to the host mailbox — no global partials, no ticket atomic, no second
// pass.  Used whenever the features fit one cluster's grid-stride budget (any VLP-16-sized sweep).
constexpr int CL_CTAS = 8, CL_NT = 256;
__device__ __forceinline__ void
    odom_iter_cluster_body(OdomT T, SinCos3 sc, int iter, const float4* __restrict__ sharp, int n_sharp, const float4* __restrict__ flat,
                             int n_flat, const float4* __restrict__ corner_last, const float4* __restrict__ surf_last,
                             const int* __restrict__ c1, const int* __restrict__ c2, const int* __restrict__ s1, const int* __restrict__ s2,
                             const int* __restrict__ s3, double* __restrict__ out28, unsigned long long seq) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ double s_part[CL_NT / 32][28];
  __shared__ double s_cta[28];
  Acc28 acc;
  acc.clear();
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  for (int q = blockIdx.x * CL_NT + tid; q < n_sharp + n_flat; q += CL_CTAS * CL_NT)
    odom_row(q, T, sc, iter, sharp, n_sharp, flat, n_flat, corner_last, surf_last, c1, c2, s1, s2, s3, acc);
  const double wsum = lg_warp_reduce28(acc.v, lane);
  if (lane < 28) s_part[w][lane] = wsum;
  __syncthreads();
  if (tid < 28) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < CL_NT / 32; k++) s += s_part[k][tid];
    s_cta[tid] = s;
  }
  cluster.sync();
  if (cluster.block_rank() == 0) {
    if (tid < 28) {
      double s = 0.0;
      for (int r = 0; r < CL_CTAS; r++) s += cluster.map_shared_rank(s_cta, r)[tid];
      out28[tid] = s;
    }
    __threadfence_system();
    __syncthreads();
    if (tid == 0) {
      *((volatile unsigned long long*)(out28 + 31)) = seq;
      __threadfence_system();
    }
  }
  cluster.sync();  // peers must not exit (and release their shared memory) before CTA 0 has read it
}


// Iterations it0 .. it1-1 of the Gauss-Newton loop in ONE launch (see lg_odom.h).  Per iteration: rows -> 31-shuffle warp
// reduction -> CTA partial in shared memory -> cluster barrier -> CTA 0 adds the partials of its peers out of their shared
// memory (DSMEM, rank order) -> thread 0 solves (Householder QR, LO:975), projects when degenerate (LO:1000-1004),
// updates the transform (LO:1006-1011 + the NaN guards LO:1013-1018), evaluates the convergence test (LO:1020-1031) and
// the sin/cos of the new angles (bit-exact libm ports) -> cluster barrier -> every CTA picks the new transform up from
// CTA 0's shared memory.  No global memory, no atomics and no host between iterations.
constexpr int LP_NT = 256;
struct OdomLoopShared {
  float T[6];
  float sc[6];  // srx crx sry cry srz crz
  int done, last_iter;
};
__device__ __forceinline__ void
    odom_loop_body(const OdomLoopArgs& A, const float4* __restrict__ sharp, int n_sharp, const float4* __restrict__ flat, int n_flat,
                     const float4* __restrict__ corner_last, const float4* __restrict__ surf_last, const int* __restrict__ c1,
                     const int* __restrict__ c2, const int* __restrict__ s1, const int* __restrict__ s2, const int* __restrict__ s3,
                     double* __restrict__ out, unsigned long long seq) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned int rank = cluster.block_rank(), nranks = cluster.num_blocks();
  __shared__ double s_part[LP_NT / 32][28];
  __shared__ double s_cta[28];
  __shared__ double s_tot[28];
  __shared__ OdomLoopShared s_state;  // authoritative copy lives in CTA 0
  __shared__ OdomLoopShared s_mine;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  OdomT T = A.T;
  SinCos3 sc = A.sc;
  int iter = A.it0;
  if (rank == 0 && tid == 0) {
    s_state.done = 0;
    s_state.last_iter = A.it0 - 1;
  }
  // one row per thread (every VLP-16-sized sweep): its inputs stay in registers for all iterations of this launch
  const bool one_row = n_sharp + n_flat <= (int)(nranks * LP_NT);
  OdomRowIn mine;
  mine.kind = 0;
  if (one_row) mine = odom_row_load(rank * LP_NT + tid, sharp, n_sharp, flat, n_flat, corner_last, surf_last, c1, c2, s1, s2, s3);
  while (true) {
    Acc28 acc;
    acc.clear();
    if (one_row) {
      odom_row_eval(mine, T, sc, iter, acc);
    } else {
      for (int q = rank * LP_NT + tid; q < n_sharp + n_flat; q += nranks * LP_NT)
        odom_row(q, T, sc, iter, sharp, n_sharp, flat, n_flat, corner_last, surf_last, c1, c2, s1, s2, s3, acc);
    }
    const double r = lg_warp_reduce28(acc.v, lane);
    if (lane < 28) s_part[w][lane] = r;
    __syncthreads();
    if (tid < 28) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < LP_NT / 32; k++) s += s_part[k][tid];
      s_cta[tid] = s;
    }
    cluster.sync();
    if (rank == 0) {
      if (tid < 28) {
        double s = 0.0;
        for (unsigned int k = 0; k < nranks; k++) s += cluster.map_shared_rank(s_cta, k)[tid];
        s_tot[tid] = s;
      }
      __syncthreads();
      if (w == 0) {  // the QR runs on seven lanes, the six sin/cos and the two convergence norms on eight, side by side
        float X[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        float AtA[36], AtB[6];
        int n_sel;
        lg_unpack28(s_tot, AtA, AtB, &n_sel);  // every lane for itself: 28 shared-memory reads
        const int solved = n_sel >= 10;        // LO:904-907
        if (solved) {
          lg_qr_solve6_warp(AtA, AtB, lane, X);
          if (A.degenerate) {
            float X2[6];
            for (int i = 0; i < 6; i++) X2[i] = X[i];
            lg_gemm_dacc(A.matP, X2, X, 6, 6, 1);
          }
#pragma unroll
          for (int i = 0; i < 6; i++) {
            float v = T.t[i] + X[i];
            if (isnan(v)) v = 0.f;
            T.t[i] = v;
          }
        }
        float Tn[6];
#pragma unroll
        for (int i = 0; i < 6; i++) Tn[i] = T.t[i];
        int small = 0;
        if (lane < 6) {
          const int k = lane >> 1;
          const float ang = k == 0 ? Tn[0] : (k == 1 ? Tn[1] : Tn[2]);
          s_state.sc[lane] = (lane & 1) ? lgm_cosf(ang) : lgm_sinf(ang);
          float tv = Tn[0];
#pragma unroll
          for (int i = 1; i < 6; i++) tv = lane == i ? Tn[i] : tv;
          s_state.T[lane] = tv;
        } else if (lane == 6) {
          const double r0 = X[0] * 180.0 / M_PI, r1 = X[1] * 180.0 / M_PI, r2 = X[2] * 180.0 / M_PI;
          small = (float)sqrt(r0 * r0 + r1 * r1 + r2 * r2) < 0.1;
        } else if (lane == 7) {
          const double t0 = X[3] * 100, t1 = X[4] * 100, t2 = X[5] * 100;
          small = (float)sqrt(t0 * t0 + t1 * t1 + t2 * t2) < 0.1;
        }
        const unsigned int both = __ballot_sync(0xffffffffu, small) & 0xc0u;
        if (lane == 0) {
          s_state.done = (solved && both == 0xc0u) ? 1 : 0;
          s_state.last_iter = iter;
        }
      }
    }
    cluster.sync();
    if (tid < (int)(sizeof(OdomLoopShared) / 4))
      reinterpret_cast<int*>(&s_mine)[tid] = reinterpret_cast<const int*>(cluster.map_shared_rank(&s_state, 0))[tid];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 6; i++) T.t[i] = s_mine.T[i];
    sc.srx = s_mine.sc[0]; sc.crx = s_mine.sc[1]; sc.sry = s_mine.sc[2]; sc.cry = s_mine.sc[3]; sc.srz = s_mine.sc[4]; sc.crz = s_mine.sc[5];
    const int done = s_mine.done;
    iter++;
    if (done || iter >= A.it1) break;
    __syncthreads();  // s_mine is rewritten next round
  }
  if (rank == 0 && tid == 0) {
#pragma unroll
    for (int i = 0; i < 6; i++) out[i] = (double)s_mine.T[i];
    out[6] = (double)s_mine.last_iter;
    out[7] = (double)s_mine.done;
    __threadfence_system();
    *((volatile unsigned long long*)(out + 31)) = seq;
  }
  cluster.sync();  // CTA 0 must outlive its peers' last read of s_state
}

// LO:156-227.  sT = sin/cos of the full transform, imu sin/cos evaluated on the host.
__device__ __forceinline__ void odom_to_end_body(OdomT T, SinCos3 sT, ImuSC imu, const float4* __restrict__ in0, float4* __restrict__ out0,
                                                           int n0, const float4* __restrict__ in1, float4* __restrict__ out1, int n1,
                                                           const float4* __restrict__ in2, float4* __restrict__ out2, int n2) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  const float4* in;
  float4* out;
  if (i < n0) {
    in = in0; out = out0;
  } else if (i < n0 + n1) {
    in = in1; out = out1; i -= n0;
  } else if (i < n0 + n1 + n2) {
    in = in2; out = out2; i -= n0 + n1;
  } else {
    return;
  }
  const float4 pi = in[i];
  const float4 a = transform_to_start(T, pi);
  float x3 = a.x, y3 = a.y, z3 = a.z;
  float x4 = sT.cry * x3 + sT.sry * z3;
  float y4 = y3;
  float z4 = -sT.sry * x3 + sT.cry * z3;
  float x5 = x4;
  float y5 = sT.crx * y4 - sT.srx * z4;
  float z5 = sT.srx * y4 + sT.crx * z4;
  float x6 = sT.crz * x5 - sT.srz * y5 + T.t[3];
  float y6 = sT.srz * x5 + sT.crz * y5 + T.t[4];
  float z6 = z5 + T.t[5];
  float x7 = imu.c_roll_s * (x6 - imu.shift[0]) - imu.s_roll_s * (y6 - imu.shift[1]);
  float y7 = imu.s_roll_s * (x6 - imu.shift[0]) + imu.c_roll_s * (y6 - imu.shift[1]);
  float z7 = z6 - imu.shift[2];
  float x8 = x7;
  float y8 = imu.c_pitch_s * y7 - imu.s_pitch_s * z7;
  float z8 = imu.s_pitch_s * y7 + imu.c_pitch_s * z7;
  float x9 = imu.c_yaw_s * x8 + imu.s_yaw_s * z8;
  float y9 = y8;
  float z9 = -imu.s_yaw_s * x8 + imu.c_yaw_s * z8;
  float x10 = imu.c_yaw_l * x9 - imu.s_yaw_l * z9;
  float y10 = y9;
  float z10 = imu.s_yaw_l * x9 + imu.c_yaw_l * z9;
  float x11 = x10;
  float y11 = imu.c_pitch_l * y10 + imu.s_pitch_l * z10;
  float z11 = -imu.s_pitch_l * y10 + imu.c_pitch_l * z10;
  float4 po;
  po.x = imu.c_roll_l * x11 + imu.s_roll_l * y11;
  po.y = -imu.s_roll_l * x11 + imu.c_roll_l * y11;
  po.z = z11;
  po.w = int(pi.w);
  out[i] = po;
}

// ------------------------------------------------------------------------------------------------ launch forms
// Like the extraction kernels (lg_extract.cu), every odometry kernel exists for ONE sequence (arguments by value) and
// BATCHED over several sequences in lock-step (grid.y = sequence, arguments from a device table of OdK, see lg_odom.h): a
// member that does not take part in a kernel of the round (converged, no refresh due) returns at once.
__global__ void __launch_bounds__(1024) odom_bounds_kernel(const float4* __restrict__ corner_last, int n_cl, const float4* __restrict__ surf_last,
                                                            int n_sl, float4* __restrict__ box_c, float4* __restrict__ sup_c,
                                                            float4* __restrict__ box_s, float4* __restrict__ sup_s) {
  odom_bounds_body(corner_last, n_cl, surf_last, n_sl, box_c, sup_c, box_s, sup_s);
}
__global__ void __launch_bounds__(1024) odom_bounds_batch_kernel(const OdK* __restrict__ tab) {
  const OdK& A = tab[blockIdx.y];
  if (!A.do_bounds || (int)blockIdx.x >= ((A.n_cl + 1023) >> 10) + ((A.n_sl + 1023) >> 10)) return;
  odom_bounds_body(A.corner_last, A.n_cl, A.surf_last, A.n_sl, A.box_c, A.sup_c, A.box_s, A.sup_s);
}
__global__ void __launch_bounds__(KP_WARPS * 32) odom_refresh_batch_kernel(const OdK* __restrict__ tab) {
  const OdK& A = tab[blockIdx.y];
  const int lane = threadIdx.x & 31;
  const int q = blockIdx.x * KP_WARPS + (threadIdx.x >> 5);
  if (!A.do_refresh || q >= A.n_sharp + A.n_flat) return;
  const OdomT T = A.do_loop ? A.la.T : A.T;
  const unsigned long long b =
      odom_knn_pruned_warp(T, q, lane, A.sharp, A.n_sharp, A.flat, A.corner_last, A.n_cl, A.surf_last, A.n_sl, A.box_c, A.sup_c, A.box_s, A.sup_s);
  if (lane == 0) A.best[q] = b;
  odom_corr_pruned_warp(T, q, lane, b, A.sharp, A.n_sharp, A.flat, A.n_flat, A.corner_last, A.n_cl, A.surf_last, A.n_sl, A.box_c, A.box_s, A.c1, A.c2,
                        A.s1, A.s2, A.s3);
}
__global__ void __cluster_dims__(CL_CTAS, 1, 1) __launch_bounds__(CL_NT)
    odom_iter_cluster_kernel(OdomT T, SinCos3 sc, int iter, const float4* __restrict__ sharp, int n_sharp, const float4* __restrict__ flat,
                             int n_flat, const float4* __restrict__ corner_last, const float4* __restrict__ surf_last,
                             const int* __restrict__ c1, const int* __restrict__ c2, const int* __restrict__ s1, const int* __restrict__ s2,
                             const int* __restrict__ s3, double* __restrict__ out28, unsigned long long seq) {
  odom_iter_cluster_body(T, sc, iter, sharp, n_sharp, flat, n_flat, corner_last, surf_last, c1, c2, s1, s2, s3, out28, seq);
}
__global__ void __cluster_dims__(CL_CTAS, 1, 1) __launch_bounds__(CL_NT) odom_iter_cluster_batch_kernel(const OdK* __restrict__ tab) {
  const OdK& A = tab[blockIdx.y];
  if (!A.do_iter0) return;  // the whole cluster of this member
  odom_iter_cluster_body(A.T, A.sc, A.iter, A.sharp, A.n_sharp, A.flat, A.n_flat, A.corner_last, A.surf_last, A.c1, A.c2, A.s1, A.s2, A.s3, A.out,
                         A.seq);
}
__global__ void __launch_bounds__(LP_NT, 2)
    odom_loop_kernel(OdomLoopArgs A, const float4* __restrict__ sharp, int n_sharp, const float4* __restrict__ flat, int n_flat,
                     const float4* __restrict__ corner_last, const float4* __restrict__ surf_last, const int* __restrict__ c1,
                     const int* __restrict__ c2, const int* __restrict__ s1, const int* __restrict__ s2, const int* __restrict__ s3,
                     double* __restrict__ out, unsigned long long seq) {
  odom_loop_body(A, sharp, n_sharp, flat, n_flat, corner_last, surf_last, c1, c2, s1, s2, s3, out, seq);
}
__global__ void __launch_bounds__(LP_NT, 2) odom_loop_batch_kernel(const OdK* __restrict__ tab) {
  const OdK& A = tab[blockIdx.y];
  if (!A.do_loop) return;
  odom_loop_body(A.la, A.sharp, A.n_sharp, A.flat, A.n_flat, A.corner_last, A.surf_last, A.c1, A.c2, A.s1, A.s2, A.s3, A.out, A.seq);
}
__global__ void __launch_bounds__(256) odom_to_end_kernel(OdomT T, SinCos3 sT, ImuSC imu, const float4* __restrict__ in0, float4* __restrict__ out0,
                                                           int n0, const float4* __restrict__ in1, float4* __restrict__ out1, int n1,
                                                           const float4* __restrict__ in2, float4* __restrict__ out2, int n2) {
  odom_to_end_body(T, sT, imu, in0, out0, n0, in1, out1, n1, in2, out2, n2);
}
__global__ void __launch_bounds__(256) odom_to_end_batch_kernel(const OdK* __restrict__ tab) {
  const OdK& A = tab[blockIdx.y];
  if (!A.do_to_end || (int)(blockIdx.x * blockDim.x) >= A.n0 + A.n1 + A.n2) return;
  odom_to_end_body(A.T, A.sT, A.imu, A.in0, A.out0, A.n0, A.in1, A.out1, A.n1, A.in2, A.out2, A.n2);
}

}  // namespace

static int odom_ensure(OdomWs& ws, int n_sharp, int n_flat, cudaStream_t st) {
  const int nq = n_sharp + n_flat;
  const int nb = std::max(1, lg_div_up(nq, IT_NT));
  LG_CHECK(ws.best.ensure((size_t)(nq + 1) * 8, st));
  LG_CHECK(ws.c1.ensure((size_t)(n_sharp + 1) * 4, st, true));
  LG_CHECK(ws.c2.ensure((size_t)(n_sharp + 1) * 4, st, true));
  LG_CHECK(ws.s1.ensure((size_t)(n_flat + 1) * 4, st, true));
  LG_CHECK(ws.s2.ensure((size_t)(n_flat + 1) * 4, st, true));
  LG_CHECK(ws.s3.ensure((size_t)(n_flat + 1) * 4, st, true));
  LG_CHECK(ws.partials.ensure((size_t)nb * 28 * 8, st));
  if (!ws.ticket.p) {
    LG_CHECK(ws.ticket.ensure(4, st));
    LG_CHECK(cudaMemsetAsync(ws.ticket.p, 0, 4, st));
  }
  return LOAM_OK;
}

// LO:603-677 / LO:758-844: nearest neighbour + ring scans for every feature (every 5th iteration)
static int odom_refresh_corr(OdomWs& ws, const OdomT& T, const float4* sharp, int n_sharp, const float4* flat, int n_flat,
                             const float4* corner_last, int n_cl, const float4* surf_last, int n_sl, cudaStream_t st, long long* launches) {
  const int nq = n_sharp + n_flat;
  if (nq <= 0) return LOAM_OK;
  static const bool brute = getenv("LOAM_ODOM_BRUTE_FORCE") != nullptr;  // cross-check path
  if (!brute && !ws.bounds_valid) {  // once per sweep: the previous sweep's clouds do not change between refreshes
    const int nsup_c = lg_div_up(n_cl, 1024), nsup_s = lg_div_up(n_sl, 1024);
    // per cloud: [2 * 32 * nsup boxes | 2 * nsup super-boxes] float4
    LG_CHECK(ws.bounds_c.ensure((size_t)(nsup_c + 1) * 66 * 16, st));
    LG_CHECK(ws.bounds_s.ensure((size_t)(nsup_s + 1) * 66 * 16, st));
    if (nsup_c + nsup_s > 0) {
      LgProfScope prof_scope(LGK_ODOM_KNN, st, 0.0);
      odom_bounds_kernel<<<nsup_c + nsup_s, 1024, 0, st>>>(corner_last, n_cl, surf_last, n_sl, ws.bounds_c.as<float4>(),
                                                           ws.bounds_c.as<float4>() + (size_t)nsup_c * 64, ws.bounds_s.as<float4>(),
                                                           ws.bounds_s.as<float4>() + (size_t)nsup_s * 64);
      (*launches)++;
    }
    ws.bounds_valid = true;
  }
  LgProfScope prof_scope(LGK_ODOM_KNN, st, (double)nq);
  if (brute) {
    LG_CHECK(cudaMemsetAsync(ws.best.p, 0xff, (size_t)nq * 8, st));
    const int tiles_c = lg_div_up(n_sharp, KNN_Q), tiles_s = lg_div_up(n_flat, KNN_Q);
    const int chunks = std::max(1, lg_div_up(std::max(n_cl, n_sl), KNN_T));
    dim3 grid(tiles_c + tiles_s, chunks);
    odom_knn_kernel<<<grid, KNN_Q, 0, st>>>(T, sharp, n_sharp, flat, n_flat, corner_last, n_cl, surf_last, n_sl, tiles_c,
                                            ws.best.as<unsigned long long>());
    odom_corr_kernel<<<lg_div_up(nq, CORR_WARPS), CORR_WARPS * 32, 0, st>>>(T, sharp, n_sharp, flat, n_flat, corner_last, n_cl, surf_last, n_sl,
                                                                            ws.best.as<unsigned long long>(), ws.c1.as<int>(), ws.c2.as<int>(),
                                                                            ws.s1.as<int>(), ws.s2.as<int>(), ws.s3.as<int>());
    (*launches) += 2;
  } else {
    const int nsup_c = lg_div_up(n_cl, 1024), nsup_s = lg_div_up(n_sl, 1024);
    odom_refresh_pruned_kernel<<<lg_div_up(nq, KP_WARPS), KP_WARPS * 32, 0, st>>>(
        T, sharp, n_sharp, flat, n_flat, corner_last, n_cl, surf_last, n_sl, ws.bounds_c.as<float4>(), ws.bounds_c.as<float4>() + (size_t)nsup_c * 64,
        ws.bounds_s.as<float4>(), ws.bounds_s.as<float4>() + (size_t)nsup_s * 64, ws.best.as<unsigned long long>(), ws.c1.as<int>(),
        ws.c2.as<int>(), ws.s1.as<int>(), ws.s2.as<int>(), ws.s3.as<int>());
    (*launches)++;
  }
  LG_CHECK(cudaGetLastError());
  return LOAM_OK;
}

int lg_odom_iter_launch(OdomWs& ws, const OdomT& T, const SinCos3& sc, int iter, const float4* sharp, int n_sharp, const float4* flat, int n_flat,
                        const float4* corner_last, int n_cl, const float4* surf_last, int n_sl, double* out28, unsigned long long seq, cudaStream_t st,
                        long long* launches) {
  const int nq = n_sharp + n_flat;
  const int nb = std::max(1, lg_div_up(nq, IT_NT));
  int rc = odom_ensure(ws, n_sharp, n_flat, st);
  if (rc) return rc;
  if (iter % 5 == 0) {
    rc = odom_refresh_corr(ws, T, sharp, n_sharp, flat, n_flat, corner_last, n_cl, surf_last, n_sl, st, launches);
    if (rc) return rc;
  }
  LgProfScope prof_scope(LGK_ODOM_ITER, st, (double)nq);
  if (nq <= CL_CTAS * CL_NT * 3) {
    odom_iter_cluster_kernel<<<CL_CTAS, CL_NT, 0, st>>>(T, sc, iter, sharp, n_sharp, flat, n_flat, corner_last, surf_last, ws.c1.as<int>(),
                                                        ws.c2.as<int>(), ws.s1.as<int>(), ws.s2.as<int>(), ws.s3.as<int>(), out28, seq);
  } else {
    odom_iter_kernel<<<nb, IT_NT, 0, st>>>(T, sc, iter, sharp, n_sharp, flat, n_flat, corner_last, n_cl, surf_last, n_sl, ws.c1.as<int>(),
                                           ws.c2.as<int>(), ws.s1.as<int>(), ws.s2.as<int>(), ws.s3.as<int>(), ws.partials.as<double>(),
                                           ws.ticket.as<unsigned int>(), out28, seq);
  }
  (*launches)++;
  LG_CHECK(cudaGetLastError());
  return LOAM_OK;
}

int lg_odom_loop_launch(OdomWs& ws, const OdomLoopArgs& args, const float4* sharp, int n_sharp, const float4* flat, int n_flat,
                        const float4* corner_last, int n_cl, const float4* surf_last, int n_sl, double* out, unsigned long long seq, cudaStream_t st,
                        long long* launches) {
  static int cluster_ctas_dev[64] = {};  // 16 where the device schedules a 16-CTA cluster, else the portable 8
  int dev = 0;
  LG_CHECK(cudaGetDevice(&dev));
  int& cluster_ctas = cluster_ctas_dev[dev & 63];
  if (cluster_ctas == 0) {
    cluster_ctas = 8;
    if (cudaFuncSetAttribute(odom_loop_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(16);
      cfg.blockDim = dim3(LP_NT);
      cudaLaunchAttribute at;
      at.id = cudaLaunchAttributeClusterDimension;
      at.val.clusterDim.x = 16; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
      cfg.attrs = &at;
      cfg.numAttrs = 1;
      int n_clusters = 0;
      if (cudaOccupancyMaxActiveClusters(&n_clusters, odom_loop_kernel, &cfg) == cudaSuccess && n_clusters > 0) cluster_ctas = 16;
    }
    (void)cudaGetLastError();
  }
  int rc = odom_ensure(ws, n_sharp, n_flat, st);
  if (rc) return rc;
  if (args.it0 % 5 == 0) {
    rc = odom_refresh_corr(ws, args.T, sharp, n_sharp, flat, n_flat, corner_last, n_cl, surf_last, n_sl, st, launches);
    if (rc) return rc;
  }
  const int nq = n_sharp + n_flat;
  LgProfScope prof_scope(LGK_ODOM_ITER, st, (double)nq * (args.it1 - args.it0));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cluster_ctas);
  cfg.blockDim = dim3(LP_NT);
  cfg.stream = st;
  cudaLaunchAttribute at;
  at.id = cudaLaunchAttributeClusterDimension;
  at.val.clusterDim.x = cluster_ctas; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
  cfg.attrs = &at;
  cfg.numAttrs = 1;
  LG_CHECK(cudaLaunchKernelEx(&cfg, odom_loop_kernel, args, sharp, n_sharp, flat, n_flat, corner_last, surf_last, (const int*)ws.c1.as<int>(),
                              (const int*)ws.c2.as<int>(), (const int*)ws.s1.as<int>(), (const int*)ws.s2.as<int>(), (const int*)ws.s3.as<int>(), out,
                              seq));
  (*launches)++;
  return LOAM_OK;
}

int lg_odom_to_end_launch(const OdomT& T, const SinCos3& sT, const ImuSC& imu, const float4* in0, float4* out0, int n0, const float4* in1,
                          float4* out1, int n1, const float4* in2, float4* out2, int n2, cudaStream_t st, long long* launches) {
  const int n = n0 + n1 + n2;
  if (n <= 0) return LOAM_OK;
  LgProfScope prof_scope(LGK_TO_END, st, (double)n);
  odom_to_end_kernel<<<lg_div_up(n, 256), 256, 0, st>>>(T, sT, imu, in0, out0, n0, in1, out1, n1, in2, out2, n2);
  (*launches)++;
  LG_CHECK(cudaGetLastError());
  return LOAM_OK;
}

// ------------------------------------------------------------------------------------------------ batched launches
int lg_odom_batch_prepare(OdomWs& ws, int n_sharp, int n_flat, int n_cl, int n_sl, cudaStream_t st, OdK* k) {
  int rc = odom_ensure(ws, n_sharp, n_flat, st);
  if (rc) return rc;
  const int nsup_c = lg_div_up(n_cl, 1024), nsup_s = lg_div_up(n_sl, 1024);
  LG_CHECK(ws.bounds_c.ensure((size_t)(nsup_c + 1) * 66 * 16, st));
  LG_CHECK(ws.bounds_s.ensure((size_t)(nsup_s + 1) * 66 * 16, st));
  k->box_c = ws.bounds_c.as<float4>(); k->sup_c = ws.bounds_c.as<float4>() + (size_t)nsup_c * 64;
  k->box_s = ws.bounds_s.as<float4>(); k->sup_s = ws.bounds_s.as<float4>() + (size_t)nsup_s * 64;
  k->best = ws.best.as<unsigned long long>();
  k->c1 = ws.c1.as<int>(); k->c2 = ws.c2.as<int>(); k->s1 = ws.s1.as<int>(); k->s2 = ws.s2.as<int>(); k->s3 = ws.s3.as<int>();
  return LOAM_OK;
}
bool lg_odom_batch_fits(int n_sharp, int n_flat) { return n_sharp + n_flat <= CL_CTAS * CL_NT * 3; }

// One lock-step round for B members: bounds -> refresh -> iteration 0 / loop block, each kernel launched once if any member
// takes part in it.  The table is copied to `tab` first (it is small: B x sizeof(OdK)).
int lg_odom_batch_round(const OdK* host_tab, int B, DevBuf& tab, cudaStream_t st, long long* launches) {
  bool any_bounds = false, any_refresh = false, any_iter0 = false, any_loop = false;
  int g_bounds = 1, g_q = 1;
  double units_r = 0, units_i = 0;
  for (int b = 0; b < B; b++) {
    const OdK& k = host_tab[b];
    any_bounds |= k.do_bounds != 0; any_refresh |= k.do_refresh != 0; any_iter0 |= k.do_iter0 != 0; any_loop |= k.do_loop != 0;
    g_bounds = std::max(g_bounds, lg_div_up(k.n_cl, 1024) + lg_div_up(k.n_sl, 1024));
    g_q = std::max(g_q, lg_div_up(k.n_sharp + k.n_flat, KP_WARPS));
    if (k.do_refresh) units_r += k.n_sharp + k.n_flat;
    if (k.do_iter0) units_i += k.n_sharp + k.n_flat;
    if (k.do_loop) units_i += (double)(k.n_sharp + k.n_flat) * (k.la.it1 - k.la.it0);
  }
  if (!(any_bounds || any_refresh || any_iter0 || any_loop)) return LOAM_OK;
  LG_CHECK(tab.ensure((size_t)B * sizeof(OdK) + 64, st));
  LG_CHECK(cudaMemcpyAsync(tab.p, host_tab, (size_t)B * sizeof(OdK), cudaMemcpyHostToDevice, st));
  const OdK* d = tab.as<OdK>();
  if (any_bounds || any_refresh) {
    LgProfScope prof_scope(LGK_ODOM_KNN, st, units_r);
    if (any_bounds) {
      odom_bounds_batch_kernel<<<dim3(g_bounds, B), 1024, 0, st>>>(d);
      (*launches)++;
    }
    if (any_refresh) {
      odom_refresh_batch_kernel<<<dim3(g_q, B), KP_WARPS * 32, 0, st>>>(d);
      (*launches)++;
    }
  }
  if (any_iter0 || any_loop) {
    LgProfScope prof_scope(LGK_ODOM_ITER, st, units_i);
    if (any_iter0) {
      odom_iter_cluster_batch_kernel<<<dim3(CL_CTAS, B), CL_NT, 0, st>>>(d);
      (*launches)++;
    }
    if (any_loop) {
      static int cluster_ctas_dev[64] = {};
      int dev = 0;
      LG_CHECK(cudaGetDevice(&dev));
      int& cluster_ctas = cluster_ctas_dev[dev & 63];
      if (cluster_ctas == 0) {
        cluster_ctas = 8;
        if (cudaFuncSetAttribute(odom_loop_batch_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
          cudaLaunchConfig_t cfg = {};
          cfg.gridDim = dim3(16);
          cfg.blockDim = dim3(LP_NT);
          cudaLaunchAttribute at;
          at.id = cudaLaunchAttributeClusterDimension;
          at.val.clusterDim.x = 16; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
          cfg.attrs = &at;
          cfg.numAttrs = 1;
          int n_clusters = 0;
          if (cudaOccupancyMaxActiveClusters(&n_clusters, odom_loop_batch_kernel, &cfg) == cudaSuccess && n_clusters > 0) cluster_ctas = 16;
        }
        (void)cudaGetLastError();
      }
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(cluster_ctas, B);
      cfg.blockDim = dim3(LP_NT);
      cfg.stream = st;
      cudaLaunchAttribute at;
      at.id = cudaLaunchAttributeClusterDimension;
      at.val.clusterDim.x = cluster_ctas; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
      cfg.attrs = &at;
      cfg.numAttrs = 1;
      LG_CHECK(cudaLaunchKernelEx(&cfg, odom_loop_batch_kernel, d));
      (*launches)++;
    }
  }
  LG_CHECK(cudaGetLastError());
  return LOAM_OK;
}

int lg_odom_batch_to_end(const OdK* host_tab, int B, DevBuf& tab, cudaStream_t st, long long* launches) {
  int g = 0;
  double units = 0;
  for (int b = 0; b < B; b++)
    if (host_tab[b].do_to_end) {
      g = std::max(g, lg_div_up(host_tab[b].n0 + host_tab[b].n1 + host_tab[b].n2, 256));
      units += host_tab[b].n0 + host_tab[b].n1 + host_tab[b].n2;
    }
  if (g == 0) return LOAM_OK;
  LG_CHECK(tab.ensure((size_t)B * sizeof(OdK) + 64, st));
  LG_CHECK(cudaMemcpyAsync(tab.p, host_tab, (size_t)B * sizeof(OdK), cudaMemcpyHostToDevice, st));
  LgProfScope prof_scope(LGK_TO_END, st, units);
  odom_to_end_batch_kernel<<<dim3(g, B), 256, 0, st>>>(tab.as<OdK>());
  (*launches)++;
  LG_CHECK(cudaGetLastError());
  return LOAM_OK;
}
