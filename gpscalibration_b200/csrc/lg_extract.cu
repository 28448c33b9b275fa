// K1-K4 — per-sweep feature extraction, the B200 replacement for the body of laserCloudHandler
// (scanRegistration.cpp SR:238-752; IMU de-skew branch SR:364-434 dormant, SURVEY §8a note).
//
//   sr_ring_kernel     SR:260-362  NaN strip, axis remap, ring table, raw azimuth, halfPassed latch as a prefix-min,
//                                  per-CTA ring histogram
//   sr_scan_kernel     SR:444-447  exclusive scan of (ring, CTA) counts -> ring-major offsets (stable bucket by ring)
//   sr_scatter_kernel  SR:340-362,436  relTime / intensity, stable scatter to the ring-major float4 cloud
//   sr_curv_kernel     SR:454-549  11-tap curvature, ring bounds, occlusion / parallel-beam conditions, gap flags
//   sr_select_kernel   SR:559-675  one CTA per ring: per-sector stable sort as ranks (counting) + the greedy pick with +-5
//                                  suppression resolved in index space (dominator rounds instead of a serial walk)
//   sr_collect_kernel  SR:587-633  compacts the picks into the four feature clouds, sets up the per-ring voxel jobs
//   (lg_vox_small)     SR:677-683  per-ring 0.2 m voxel grid of the less-flat points
//   sr_concat_kernel   SR:683      concatenates the per-ring results
// Every array is ring-major SoA/float4; algorithmic bytes per sweep 28 N + 16 F.
#include "lg_extract.h"

namespace {

constexpr int SR_NT = 256;
// Points per thread of the ring / scatter kernels: 1 for sweeps up to 64 k points (a 28.8 k-point sweep then spreads over 113
// CTAs instead of 29: the per-point atanf / atan2f chains are the kernel's latency), 4 above (keeps the (ring, CTA) table
// that one CTA scans short).  The kernels derive it from n and the CTA count, so both see the same tiling.
static inline int sr_items_for(int n) { return n <= 65536 ? 1 : 4; }
__device__ __forceinline__ int sr_items(int n, int nblocks) { return (n + nblocks * SR_NT - 1) / (nblocks * SR_NT); }
constexpr int MAXR = 64;

// cond bits written by sr_curv_kernel
constexpr unsigned char C_A = 1;    // SR:508-520 marks i-5..i
constexpr unsigned char C_B = 2;    // SR:521-533 marks i+1..i+6
constexpr unsigned char C_C = 4;    // SR:546-548 marks i
constexpr unsigned char C_GAP = 8;  // gap^2(p_i, p_{i-1}) > 0.05 (SR:604,617,648,661)

__device__ __forceinline__ bool finite3(const float* p) { return isfinite(p[0]) && isfinite(p[1]) && isfinite(p[2]); }
__device__ __forceinline__ const float* pt_at(const float* xyz, int stride_bytes, int i) {
  return (const float*)((const char*)xyz + (size_t)i * stride_bytes);
}

// SR:297-320
__device__ __forceinline__ int ring_of(const SrParams& prm, float angle) {
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

// atan / atan2 exactly as the host libm (glibc atanf / atan2f) returns them — see lg_libm.cuh.
__device__ __forceinline__ float atanf_cr(float a) { return lgm_atanf(a); }
__device__ __forceinline__ float atan2f_cr(float y, float x) { return lgm_atan2f(y, x); }

// SR:265-278: start / end azimuth from the first / last finite point.  Evaluated by one thread per CTA.
__device__ void sweep_ori(const float* xyz, int n, int stride_bytes, float* startOri, float* endOri, int* ok) {
  int f = 0;
  while (f < n && !finite3(pt_at(xyz, stride_bytes, f))) f++;
  if (f >= n) {
    *ok = 0;
    *startOri = 0.f;
    *endOri = 0.f;
    return;
  }
  int l = n - 1;
  while (l > f && !finite3(pt_at(xyz, stride_bytes, l))) l--;
  const float* pf = pt_at(xyz, stride_bytes, f);
  const float* pl = pt_at(xyz, stride_bytes, l);
  float so = -atan2f_cr(pf[1], pf[0]);
  float eo = (float)(-atan2f_cr(pl[1], pl[0]) + 2 * M_PI);
  if (eo - so > 3 * M_PI) {
    eo = (float)(eo - 2 * M_PI);
  } else if (eo - so < M_PI) {
    eo = (float)(eo + 2 * M_PI);
  }
  *startOri = so;
  *endOri = eo;
  *ok = 1;
}

__device__ __forceinline__ void sr_ring_kernel_body(SrParams prm, const float* __restrict__ xyz, int n, int stride_bytes,
                                                         signed char* __restrict__ ring8, float* __restrict__ ori_raw,
                                                         unsigned int* __restrict__ hist, int nblocks, int* __restrict__ meta) {
  __shared__ unsigned int h[MAXR];
  __shared__ float s_ori[2];
  __shared__ int s_ok;
  __shared__ int s_jmin;
  const int tid = threadIdx.x;
  if (tid < MAXR) h[tid] = 0;
  if (tid == 0) {
    sweep_ori(xyz, n, stride_bytes, &s_ori[0], &s_ori[1], &s_ok);
    s_jmin = 0x7fffffff;
  }
  __syncthreads();
  const float startOri = s_ori[0];
  int jmin = 0x7fffffff;
  const int items = sr_items(n, nblocks);
  const int base = blockIdx.x * items * SR_NT;
  for (int c = 0; c < items; c++) {
    int i = base + c * SR_NT + tid;
    if (i >= n) break;
    const float* p = pt_at(xyz, stride_bytes, i);
    int ring = -1;
    float ori = 0.f;
    if (s_ok && finite3(p)) {
      float px = p[1], py = p[2], pz = p[0];  // SR:293-295
      float angle = (float)(atanf_cr(py / sqrtf(px * px + pz * pz)) * 180 / M_PI);
      ring = ring_of(prm, angle);
      if (ring >= 0) {
        ori = -atan2f_cr(px, pz);
        // SR:341-350 evaluated with the first-branch formula: the latch sets at the first kept point where it holds
        float o1 = ori;
        if (o1 < startOri - M_PI / 2) {
          o1 = (float)(o1 + 2 * M_PI);
        } else if (o1 > startOri + M_PI * 3 / 2) {
          o1 = (float)(o1 - 2 * M_PI);
        }
        if (o1 - startOri > M_PI) jmin = min(jmin, i);
        atomicAdd(&h[ring], 1u);
      }
    }
    ring8[i] = (signed char)ring;
    ori_raw[i] = ori;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) jmin = min(jmin, __shfl_xor_sync(0xffffffffu, jmin, o));
  if ((tid & 31) == 0 && jmin != 0x7fffffff) atomicMin(&s_jmin, jmin);
  __syncthreads();
  if (tid == 0 && s_jmin != 0x7fffffff) atomicMin(&meta[SRM_JSTAR], s_jmin);
  if (tid < prm.n_scans) hist[tid * nblocks + blockIdx.x] = h[tid];
}

__device__ __forceinline__ void sr_scan_kernel_body(SrParams prm, unsigned int* __restrict__ hist, int nblocks, int* __restrict__ meta) {
  __shared__ int s_scan[34];
  const int total = prm.n_scans * nblocks;
  int per = (total + 1023) / 1024;
  int b = min((int)threadIdx.x * per, total), e = min(b + per, total);
  int local = 0;
  for (int i = b; i < e; i++) local += (int)hist[i];
  int tot;
  int r = block_excl_scan<1024>(local, &tot, s_scan);
  for (int i = b; i < e; i++) {
    int v = (int)hist[i];
    hist[i] = (unsigned int)r;
    r += v;
  }
  __syncthreads();
  if (threadIdx.x < prm.n_scans) {
    meta[SRM_RING_START + threadIdx.x] = (int)hist[threadIdx.x * nblocks];
    meta[SRM_SCAN_START + threadIdx.x] = 0;  // SR:251-253 std::vector<int>(N_SCANS, 0)
    meta[SRM_SCAN_END + threadIdx.x] = 0;
  }
  if (threadIdx.x == 0) {
    meta[SRM_RING_START + prm.n_scans] = tot;
    meta[SRM_N_FULL] = tot;
    meta[SRM_VOX_OVERFLOW] = 0;
    meta[SRM_ERR] = 0;
    meta[SRM_VIRTUAL] = 0;
    int empty = 0;
    for (int r = 0; r < prm.n_scans; r++)
      if (hist[r * nblocks] == ((r + 1 < prm.n_scans) ? hist[(r + 1) * nblocks] : (unsigned int)tot)) empty = 1;
    meta[SRM_EMPTY_RING] = empty;
  }
}

__device__ __forceinline__ void sr_scatter_kernel_body(SrParams prm, const float* __restrict__ xyz, int n, int stride_bytes,
                                                            const signed char* __restrict__ ring8, const float* __restrict__ ori_raw,
                                                            const unsigned int* __restrict__ hist, int nblocks, const int* __restrict__ meta,
                                                            float4* __restrict__ full, const float* __restrict__ imu_pts = nullptr) {
  __shared__ unsigned int s_base[MAXR];
  __shared__ unsigned int s_wcnt[SR_NT / 32][MAXR];
  __shared__ float s_ori[2];
  __shared__ int s_ok;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid < MAXR) s_base[tid] = tid < prm.n_scans ? hist[tid * nblocks + blockIdx.x] : 0u;
  if (tid == 0) sweep_ori(xyz, n, stride_bytes, &s_ori[0], &s_ori[1], &s_ok);
  __syncthreads();
  const float startOri = s_ori[0], endOri = s_ori[1];
  const int jstar = meta[SRM_JSTAR];
  const int items = sr_items(n, nblocks);
  const int base = blockIdx.x * items * SR_NT;
  for (int c = 0; c < items; c++) {
    if (tid < MAXR) {
#pragma unroll
      for (int k = 0; k < SR_NT / 32; k++) s_wcnt[k][tid] = 0;
    }
    __syncthreads();
    int i = base + c * SR_NT + tid;
    int ring = (i < n) ? (int)ring8[i] : -1;
    bool valid = ring >= 0;
    unsigned int m = __match_any_sync(0xffffffffu, valid ? (unsigned int)ring : (256u + lane));
    unsigned int rank = __popc(m & ((1u << lane) - 1u));
    if (valid && rank == 0) s_wcnt[w][ring] = __popc(m);
    __syncthreads();
    if (valid) {
      unsigned int off = s_base[ring] + rank;
      for (int k = 0; k < w; k++) off += s_wcnt[k][ring];
      const float* p = pt_at(xyz, stride_bytes, i);
      float ori = ori_raw[i];
      if (i <= jstar) {  // halfPassed still false when this point is handled (SR:341-350)
        if (ori < startOri - M_PI / 2) {
          ori = (float)(ori + 2 * M_PI);
        } else if (ori > startOri + M_PI * 3 / 2) {
          ori = (float)(ori - 2 * M_PI);
        }
      } else {  // SR:351-359
        ori = (float)(ori + 2 * M_PI);
        if (ori < endOri - M_PI * 3 / 2) {
          ori = (float)(ori + 2 * M_PI);
        } else if (ori > endOri + M_PI / 2) {
          ori = (float)(ori - 2 * M_PI);
        }
      }
      float relTime = (ori - startOri) / (endOri - startOri);
      float inten = (float)(ring + prm.scan_period * relTime);  // SR:362, scanPeriod is a double
      if (imu_pts) full[off] = make_float4(imu_pts[3 * (size_t)i], imu_pts[3 * (size_t)i + 1], imu_pts[3 * (size_t)i + 2], inten);  // SR:364-434 ran
      else full[off] = make_float4(p[1], p[2], p[0], inten);
    }
    __syncthreads();
    if (tid < MAXR) {
      unsigned int add = 0;
#pragma unroll
      for (int k = 0; k < SR_NT / 32; k++) add += s_wcnt[k][tid];
      s_base[tid] += add;
    }
    __syncthreads();
  }
}

__device__ __forceinline__ float gap2(float4 a, float4 b) {
  float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
  return dx * dx + dy * dy + dz * dz;
}

__device__ __forceinline__ void sr_curv_kernel_body(SrParams prm, const float4* __restrict__ c, int* __restrict__ meta,
                                                       float* __restrict__ curv, unsigned char* __restrict__ cond, signed char* __restrict__ label) {
  const int n = meta[SRM_N_FULL];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = c[i];
  unsigned char cf = 0;
  float cv = 0.f;
  if (i >= 1 && gap2(p, c[i - 1]) > 0.05) cf |= C_GAP;
  if (i >= 5 && i < n - 5) {  // SR:454-488
    float4 m5 = c[i - 5], m4 = c[i - 4], m3 = c[i - 3], m2 = c[i - 2], m1 = c[i - 1];
    float4 q1 = c[i + 1], q2 = c[i + 2], q3 = c[i + 3], q4 = c[i + 4], q5 = c[i + 5];
    float dX = m5.x + m4.x + m3.x + m2.x + m1.x - 10 * p.x + q1.x + q2.x + q3.x + q4.x + q5.x;
    float dY = m5.y + m4.y + m3.y + m2.y + m1.y - 10 * p.y + q1.y + q2.y + q3.y + q4.y + q5.y;
    float dZ = m5.z + m4.z + m3.z + m2.z + m1.z - 10 * p.z + q1.z + q2.z + q3.z + q4.z + q5.z;
    cv = dX * dX + dY * dY + dZ * dZ;
    int sc = int(p.w);
    int prev = (i == 5) ? -1 : int(m1.w);
    if (sc != prev && sc > 0 && sc < prm.n_scans) {  // last writer wins in the reference == largest i
      atomicMax(&meta[SRM_SCAN_START + sc], i + 5);
      atomicMax(&meta[SRM_SCAN_END + sc - 1], i - 5);
    }
    if (i < n - 6) {  // SR:492-549
      float diff = gap2(q1, p);
      if (diff > 0.1) {
        float depth1 = sqrtf(p.x * p.x + p.y * p.y + p.z * p.z);
        float depth2 = sqrtf(q1.x * q1.x + q1.y * q1.y + q1.z * q1.z);
        if (depth1 > depth2) {
          float dx = q1.x - p.x * depth2 / depth1;
          float dy = q1.y - p.y * depth2 / depth1;
          float dz = q1.z - p.z * depth2 / depth1;
          if (sqrtf(dx * dx + dy * dy + dz * dz) / depth2 < 0.1) cf |= C_A;
        } else {
          float dx = q1.x * depth1 / depth2 - p.x;
          float dy = q1.y * depth1 / depth2 - p.y;
          float dz = q1.z * depth1 / depth2 - p.z;
          if (sqrtf(dx * dx + dy * dy + dz * dz) / depth1 < 0.1) cf |= C_B;
        }
      }
      float diff2 = gap2(p, m1);
      float dis = p.x * p.x + p.y * p.y + p.z * p.z;
      if (diff > 0.0002 * dis && diff2 > 0.0002 * dis) cf |= C_C;
    }
  }
  curv[i] = cv;
  cond[i] = cf;
  label[i] = 0;
}

// cloudNeighborPicked after SR:492-549 in gather form (the reference scatters; OR is order-independent).
__device__ __forceinline__ unsigned char mask_at(const unsigned char* __restrict__ cond, int i, int n) {
  unsigned char r = (cond[i] & C_C) ? 1 : 0;
#pragma unroll
  for (int j = 0; j <= 5; j++)
    if (i + j < n && (cond[i + j] & C_A)) r = 1;
#pragma unroll
  for (int j = 1; j <= 6; j++)
    if (i - j >= 0 && (cond[i - j] & C_B)) r = 1;
  return r;
}

constexpr int SEL_CAP = 2048;   // points per sector the shared-memory sort accepts
constexpr int RING_CAP = 8192;  // points per ring whose flags are staged in shared memory

// SR:597-622 / 641-666: how far the suppression of a pick at `ind` reaches forwards / backwards (0..5 each), from the
// gap flags alone — independent of the pick state, which is what lets a warp resolve 32 candidates at once.  Evaluated
// once per ring point (all threads) and packed as nf | nb << 4.
// FENCE (iv) (bounds) as in oracle/orc_sr.h suppress_neighbours.
__device__ __forceinline__ unsigned char suppress_reach(const unsigned char* s_cond, int li, int len) {
  int nf = 0;
  for (int l = 1; l <= 5; l++) {
    if (li + l >= len) break;
    if (s_cond[li + l] & C_GAP) break;
    nf = l;
  }
  int nb = 0;
  for (int l = 1; l <= 5; l++) {
    if (li - l < 0) break;
    if (s_cond[li - l + 1] & C_GAP) break;
    nb = l;
  }
  return (unsigned char)(nf | (nb << 4));
}

constexpr int SEL_NT = 512;
constexpr int SEL_WORDS = SEL_CAP / 32 + 2;  // candidate bit sets of one sector, one guard word on either side
// dynamic shared memory of sr_select_kernel: curvature key (4 B), rank / walk order / dominator mask (2 B each) per ring point
constexpr int SEL_SMEM = RING_CAP * (4 + 2 + 2 + 2);
#ifdef LG_SEL_DEBUG  // phase stamps of ring 1's CTA (tools/probe/sel_time.py): cycles in setup / ranks / dominators / rounds / numbering
__device__ long long g_sel_dbg[8];
#define SEL_STAMP(k) do { const long long t_ = clock64(); dbg_acc[k] += t_ - dbg_t; dbg_t = t_; } while (0)
#define SEL_DBG_ARGS dbg_acc, dbg_t
#else
#define SEL_STAMP(k) do { } while (0)
#define SEL_DBG_ARGS nullptr, sel_dbg_dummy
#endif

// SR:568-576 as ranks: the position of every point in its sector's stable sort = the number of keys (curvature bits,
// index) below its own.  One work item = four points of a sector (they share the key stream).  The fast form compares the
// 31-bit curvature keys alone (one subtract, one shift-accumulate per pair); equal curvatures make ranks collide, which the
// caller detects (an order slot stays empty) and then repeats the pass with EXACT = true (index as tie-break).
template <bool EXACT>
__device__ __forceinline__ void sel_rank_pass(const unsigned int* s_key, unsigned short* s_rank, unsigned short* s_sorted,
                                              const int* s_spl, const int* s_m, const int* s_qoff) {
  for (int item = threadIdx.x; item < s_qoff[6]; item += SEL_NT) {
    int j = 0;
    while (item >= s_qoff[j + 1]) j++;
    const int m = s_m[j], spl = s_spl[j], Q = (m + 3) >> 2, q = item - s_qoff[j];
    const unsigned int* sk = s_key + spl;
    const int t0 = q, t1 = q + Q, t2 = q + 2 * Q, t3 = q + 3 * Q;
    const unsigned int k0 = sk[t0], k1 = t1 < m ? sk[t1] : 0u, k2 = t2 < m ? sk[t2] : 0u, k3 = t3 < m ? sk[t3] : 0u;
    int r0 = 0, r1 = 0, r2 = 0, r3 = 0;
#pragma unroll 8
    for (int t = 0; t < m; t++) {
      const unsigned int k = sk[t];
      if (EXACT) {
        r0 += (k < k0) || (k == k0 && t < t0);
        r1 += (k < k1) || (k == k1 && t < t1);
        r2 += (k < k2) || (k == k2 && t < t2);
        r3 += (k < k3) || (k == k3 && t < t3);
      } else {  // keys are below 2^31: the sign of the difference is the comparison
        r0 += (k - k0) >> 31;
        r1 += (k - k1) >> 31;
        r2 += (k - k2) >> 31;
        r3 += (k - k3) >> 31;
      }
    }
    s_rank[spl + t0] = (unsigned short)r0;
    s_sorted[spl + r0] = (unsigned short)(spl + t0);
    if (t1 < m) s_rank[spl + t1] = (unsigned short)r1, s_sorted[spl + r1] = (unsigned short)(spl + t1);
    if (t2 < m) s_rank[spl + t2] = (unsigned short)r2, s_sorted[spl + r2] = (unsigned short)(spl + t2);
    if (t3 < m) s_rank[spl + t3] = (unsigned short)r3, s_sorted[spl + r3] = (unsigned short)(spl + t3);
  }
}

// Named barrier of the warps that run the walks (the others wait at the CTA barrier behind them).
__device__ __forceinline__ void sel_bar(int nthr) { asm volatile("barrier.sync 1, %0;" ::"r"(nthr) : "memory"); }
__device__ __forceinline__ bool sel_bar_or(int nthr, bool pred) {
  int r;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\tsetp.ne.u32 q, %1, 0;\n\tbarrier.red.or.pred p, 1, %2, q;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(r)
      : "r"((int)pred), "r"(nthr)
      : "memory");
  return r != 0;
}

// The twelve walks of a ring (six sectors, sharp then flat) on bit sets; NCH = sector points per thread (1: sectors of up
// to SEL_NT points, the usual case; 4: up to SEL_CAP).  Run by the first nthr / 32 warps of the CTA.  See the kernel's
// header for the scheme.  s_sb: IN candidates by walk position (set when a candidate turns IN) -> pick numbers by popcount.
template <int NCH>
__device__ __forceinline__ void sel_walks(const unsigned short* s_rank, const unsigned short* s_dom, const unsigned char* s_reach,
                                          unsigned char* s_picked, signed char* s_label, unsigned int (*s_bits)[2][SEL_WORDS],
                                          unsigned int* s_sb, int (*s_base)[3], const int* s_spl, const int* s_m, int a, int* my_picks,
                                          int nthr, int* phases_out, long long* dbg_acc, long long& dbg_t) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int phase = 0;
  for (int j = 0; j < 6; j++) {
    const int m = s_m[j];
    if (m <= 0) continue;
    const int spl = s_spl[j], nw = (m + 31) >> 5;
    unsigned int d_lo[NCH], d_mid[NCH], d_hi[NCH];  // the point's dominators as masks of its own and the two adjacent words
    int rk[NCH];
    unsigned int kinds[NCH];
#pragma unroll
    for (int c = 0; c < NCH; c++) {
      const int t = c * SEL_NT + tid;
      const unsigned int dm = t < m ? s_dom[spl + t] : 0u;
      rk[c] = t < m ? s_rank[spl + t] : 0;
      kinds[c] = dm >> 11;
      const unsigned long long x = (unsigned long long)(dm & 0x7ffu) << (lane + 27);  // window bit k <-> position lane + k - 5
      d_lo[c] = (unsigned int)x;
      d_mid[c] = (unsigned int)(x >> 32);
      d_hi[c] = lane >= 27 ? (dm & 0x7ffu) >> (37 - lane) : 0u;
    }
    for (int kind = 0; kind < 2; kind++, phase++) {  // 0: SR:578-624 from the largest curvature down, 1: SR:626-668 from the smallest up
      int cur = 0;
      bool have = false;
      int pos[NCH];
      bool is_in[NCH];
#pragma unroll
      for (int c = 0; c < NCH; c++) {  // candidates of this walk that no earlier pick has marked
        const int t = c * SEL_NT + tid;
        pos[c] = kind == 0 ? m - 1 - rk[c] : rk[c];
        is_in[c] = false;
        if (c * SEL_NT + (warp << 5) < m) {
          const bool cand = t < m && ((kinds[c] >> kind) & 1u) && s_picked[spl + t] == 0;
          const unsigned int bal = __ballot_sync(0xffffffffu, cand);
          if (lane == 0) {
            s_bits[0][0][(t >> 5) + 1] = 0u;
            s_bits[0][1][(t >> 5) + 1] = bal;
          }
          have |= bal != 0u;
        }
      }
      if (tid < 8) s_bits[tid >> 2][(tid >> 1) & 1][(tid & 1) ? nw + 1 : 0] = 0u;  // guard words of both buffers
      for (int w = tid; w < (NCH == 1 ? SEL_NT / 32 : nw); w += nthr) s_sb[w] = 0u;
      const bool any_cand = sel_bar_or(nthr, have);
      SEL_STAMP(2);
      if (!any_cand) {
        if (tid == 0) s_base[phase + 1][0] = s_base[phase][0], s_base[phase + 1][1] = s_base[phase][1], s_base[phase + 1][2] = s_base[phase][2];
        continue;
      }
      for (;;) {  // one round: every undecided candidate looks at its dominators' state of the previous round
        bool und = false;
#pragma unroll
        for (int c = 0; c < NCH; c++) {
          if (c * SEL_NT + (warp << 5) >= m) continue;  // warp-uniform
          const int w = c * (SEL_NT / 32) + warp + 1;
          const unsigned int i_lo = s_bits[cur][0][w - 1], i_hi = s_bits[cur][0][w + 1];
          const unsigned int u_lo = s_bits[cur][1][w - 1], u_hi = s_bits[cur][1][w + 1];
          unsigned int i_mid = s_bits[cur][0][w], u_mid = s_bits[cur][1][w];
          // the warp's own 32 points may resolve several steps of a chain on what it already knows (the neighbour words
          // keep the previous round's state, which is only ever less decided)
          for (int sub = 0; sub < 4; sub++) {
            const bool mine = (u_mid >> lane) & 1u;
            const unsigned int in_hit = (i_lo & d_lo[c]) | (i_mid & d_mid[c]) | (i_hi & d_hi[c]);
            const unsigned int un_hit = (u_lo & d_lo[c]) | (u_mid & d_mid[c]) | (u_hi & d_hi[c]);
            const bool to_out = mine && in_hit != 0u;
            const bool to_in = mine && (in_hit | un_hit) == 0u;
            const unsigned int b_in = __ballot_sync(0xffffffffu, to_in), b_out = __ballot_sync(0xffffffffu, to_out);
            if (to_in) {
              is_in[c] = true;
              atomicOr(&s_sb[pos[c] >> 5], 1u << (pos[c] & 31));
            }
            i_mid |= b_in;
            u_mid &= ~(b_in | b_out);
            if (u_mid == 0u || (b_in | b_out) == 0u) break;  // warp-uniform
          }
          const unsigned int n_und = u_mid;
          if (lane == 0) {
            s_bits[cur ^ 1][0][w] = i_mid;
            s_bits[cur ^ 1][1][w] = n_und;
          }
          und |= n_und != 0u;
        }
        cur ^= 1;
#ifdef LG_SEL_DEBUG
        dbg_acc[6]++;
#endif
        if (!sel_bar_or(nthr, und)) break;
      }
      SEL_STAMP(3);
      // the pick number = position among the IN candidates in walk order; the count limit keeps a prefix, and only kept
      // picks leave marks
      const int keep = kind == 0 ? 20 : 32, marking = kind == 0 ? 20 : 31;
      const int base_sharp = s_base[phase][0], base_less = s_base[phase][1], base_flat = s_base[phase][2];
      int total = 0;
#pragma unroll
      for (int c = 0; c < NCH; c++) {
        int num = 1;
        const int pw = pos[c] >> 5;
        const unsigned int low = (1u << (pos[c] & 31)) - 1u;
        if (NCH == 1) {  // sixteen words at most: independent loads, no serial loop (words beyond the sector are zero)
          if (__any_sync(0xffffffffu, is_in[c]) || tid == 0) {
#pragma unroll
            for (int w = 0; w < SEL_NT / 32; w++) {
              const unsigned int x = s_sb[w];
              total += __popc(x);
              num += __popc(x & (w < pw ? 0xffffffffu : (w == pw ? low : 0u)));
            }
          }
        } else {
          if (is_in[c]) {
            num += __popc(s_sb[pw] & low);
            for (int w = 0; w < pw; w++) num += __popc(s_sb[w]);
          }
          if (tid == 0 && c == 0)
            for (int w = 0; w < nw; w++) total += __popc(s_sb[w]);
        }
        if (!is_in[c] || num > keep) continue;
        const int li = spl + c * SEL_NT + tid;
        if (kind == 0) {
          if (num <= 16) {
            s_label[li] = 2;
            my_picks[SR_PICK_SHARP + base_sharp + num - 1] = li + a;
          } else {
            s_label[li] = 1;
          }
          my_picks[SR_PICK_LESS + base_less + num - 1] = li + a;
        } else {
          s_label[li] = -1;
          my_picks[SR_PICK_FLAT + base_flat + num - 1] = li + a;
        }
        if (num <= marking) {  // SR:635-638: the 32nd flat point is kept but neither marked nor suppressing
          const unsigned char rr = s_reach[li];
          const int nf = rr & 15, nb = rr >> 4;
#pragma unroll
          for (int l = -5; l <= 5; l++)
            if (l >= -nb && l <= nf) s_picked[li + l] = 1;
        }
      }
      if (tid == 0) {
        const int npick = min(total, keep);  // SR:592-594: the 21st sharp candidate only ends the walk
        s_base[phase + 1][0] = base_sharp + (kind == 0 ? min(npick, 16) : 0);
        s_base[phase + 1][1] = base_less + (kind == 0 ? npick : 0);
        s_base[phase + 1][2] = base_flat + (kind == 0 ? 0 : npick);
      }
      sel_bar(nthr);
      SEL_STAMP(4);
    }
  }
  if (tid == 0) *phases_out = phase;
}

// One CTA per ring, no serial walk.  The reference sorts a sector by curvature and then walks it greedily ("take the
// next candidate unless an earlier pick suppressed it", SR:578-668).  Both steps are restated in INDEX space:
//   sort  -> ranks by counting (sel_rank_pass): no barriers, no exchanges;
//   walk  -> a pick only suppresses points within +-5 of it, so whether a candidate is picked depends only on the
//            candidates within +-5 that come EARLIER in the walk (better rank) and whose suppression span covers it: its
//            "dominators", an 11-bit window mask fixed per point (a point is a candidate of exactly one of the two walks:
//            curvature > 0.1 or < 0.1).  A candidate is picked iff no dominator is picked.  All candidates of a sector
//            resolve this together in rounds on two bit sets (IN, UNDECIDED; one word per warp, neighbours read through
//            a funnel shift): UNDECIDED -> IN once no dominator is IN or UNDECIDED, -> OUT as soon as one is IN.  The
//            best-ranked undecided candidate always resolves; 3-4 rounds are typical.  By induction on the rank the result
//            is the walk's pick set had it no count limit; the limit (16 + 4 sharp, 32 flat, SR:592-594, 635-638) cuts
//            that set in walk order -- a prefix, since a pick never depends on later ones -- and only the kept picks
//            leave marks for the next phase (flat after sharp, sector after sector, SR:561-668).
__device__ __forceinline__ void sr_select_kernel_body(SrParams prm, const float4* __restrict__ c, int* __restrict__ meta,
                                                            const float* __restrict__ curv, const unsigned char* __restrict__ cond,
                                                            unsigned char* __restrict__ picked, unsigned char* __restrict__ mask_diag,
                                                            signed char* __restrict__ label, int* __restrict__ picks,
                                                            int* __restrict__ sort_ind, unsigned char* __restrict__ stale) {
  extern __shared__ unsigned int s_key[];  // [RING_CAP] curvature bits (clamped below 2^31)
  unsigned short* s_rank = reinterpret_cast<unsigned short*>(s_key + RING_CAP);  // position of the point in its sector's sort
  unsigned short* s_sorted = s_rank + RING_CAP;                                   // sector start + position -> local index
  unsigned short* s_dom = s_sorted + RING_CAP;  // bit 5-d: li-d dominates, bit 5+d: li+d; bit 11 / 12: candidate of the sharp / flat walk
  __shared__ unsigned char s_cond_raw[RING_CAP + 16];
  __shared__ unsigned char s_reach[RING_CAP];
  __shared__ unsigned char s_picked[RING_CAP];
  __shared__ signed char s_label[RING_CAP];
  __shared__ unsigned int s_bits[2][2][SEL_WORDS];  // [buffer][IN, UNDECIDED][guard + word of 32 sector points]
  __shared__ unsigned int s_sb[SEL_WORDS];  // IN candidates of the current walk by walk position
  __shared__ int s_spl[6], s_m[6], s_qoff[7];
  __shared__ int s_base[13][3];  // picks (sharp, less sharp, flat) of the ring before each of the twelve walks
  __shared__ int s_phases;
  unsigned char* s_cond = s_cond_raw + 8;  // [-8, len + 8): the flags of the points around the ring's range as well
  const int r = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#ifdef LG_SEL_DEBUG
  long long dbg_acc[7] = {0, 0, 0, 0, 0, 0, 0};
  long long dbg_t = clock64();
#endif
  const int n = meta[SRM_N_FULL];
  const int R = prm.n_scans;
  const int S = (r == 0) ? 5 : meta[SRM_SCAN_START + r];         // SR:489
  const int E = (r == R - 1) ? n - 5 : meta[SRM_SCAN_END + r];  // SR:490
  int* my_picks = picks + r * SR_PICKS_PER_RING;
  int* cnt = meta + SRM_PICK_CNT + r * 3;
  // SR:480-490 with empty rings: scanStartInd[r] is only written when the loop SEES ring r begin, scanEndInd[r] only when it
  // sees ring r+1 begin.  A ring whose start was never written keeps 0 and -- if its end was -- spans [0, E): a VIRTUAL
  // ring that overlaps every earlier ring (true VLP-16 angles: rings 6, 8, 10 of the reference's table).  Those depend on
  // the picks of the rings before them and are replayed in ring order by sr_virtual_kernel afterwards; here every CTA
  // only leaves the state that kernel needs (pick flags, sort order) in global memory when the sweep has one.
  int vflag = 0;
  if (tid > 0 && tid < R) {
    const int e_t = (tid == R - 1) ? n - 5 : meta[SRM_SCAN_END + tid];
    vflag = meta[SRM_SCAN_START + tid] == 0 && e_t > 0;
  }
  const bool any_virtual = __syncthreads_or(vflag) != 0;
  const bool is_virtual = r > 0 && S == 0 && E > 0;
  if (any_virtual) {  // this ring's own points: state as SR:454-549 leaves it
    const int p0 = meta[SRM_RING_START + r], p1 = meta[SRM_RING_START + r + 1];
    for (int i = p0 + tid; i < p1; i += SEL_NT) {
      const unsigned char m = mask_at(cond, i, n);
      picked[i] = m;
      mask_diag[i] = m;
      sort_ind[i] = i;
    }
    if (r == 0 && tid == 0) atomicExch(&meta[SRM_VIRTUAL], 1);
  }
  // this ring's points: [S - 5, E + 5)
  const int a = max(S - 5, 0), b = min(E + 5, n);
  const int len = b - a;
  if (is_virtual || len <= 0) {  // virtual: replayed later; empty range (end never written): the reference's loops do not run
    if (tid == 0) cnt[0] = cnt[1] = cnt[2] = 0;
    return;
  }
  if (len > RING_CAP) {
    if (tid == 0) {
      cnt[0] = cnt[1] = cnt[2] = 0;
      atomicExch(&meta[SRM_ERR], 1);
    }
    return;
  }
  for (int x = tid - 8; x < len + 8; x += SEL_NT) {  // one global read per flag; the masks below come from shared memory
    const int gi = a + x;
    s_cond[x] = (gi >= 0 && gi < n) ? cond[gi] : (unsigned char)0;
  }
  for (int li = tid; li < len; li += SEL_NT) {
    s_key[li] = min(__float_as_uint(curv[a + li]), 0x7fffffffu);
    s_sorted[li] = 0xffffu;
    s_label[li] = 0;
  }
  // sector bounds (SR:561-562); a sector that does not fit is skipped and flagged, as before
  int sp6[6], m6[6];
  bool too_big = false;
#pragma unroll
  for (int j = 0; j < 6; j++) {
    const int sp = (S * (6 - j) + E * j) / 6;
    const int ep = (S * (5 - j) + E * (j + 1)) / 6 - 1;
    int m = ep - sp + 1;
    if (m > 0 && (m > SEL_CAP || sp < a || ep >= b)) {
      too_big = true;
      m = 0;
    }
    sp6[j] = sp;
    m6[j] = max(m, 0);
  }
  if (tid == 0) {
    int q = 0;
#pragma unroll
    for (int j = 0; j < 6; j++) {
      s_spl[j] = sp6[j] - a;
      s_m[j] = m6[j];
      s_qoff[j] = q;
      q += (m6[j] + 3) >> 2;
    }
    s_qoff[6] = q;
    s_base[0][0] = s_base[0][1] = s_base[0][2] = 0;
    s_phases = 0;
  }
  __syncthreads();
  for (int li = tid; li < len; li += SEL_NT) {  // cloudNeighborPicked after SR:492-549 (mask_at, from the staged flags)
    unsigned char m = (s_cond[li] & C_C) ? 1 : 0;
#pragma unroll
    for (int j = 0; j <= 5; j++) m |= (s_cond[li + j] & C_A) ? 1 : 0;
#pragma unroll
    for (int j = 1; j <= 6; j++) m |= (s_cond[li - j] & C_B) ? 1 : 0;
    s_picked[li] = m;
    mask_diag[a + li] = m;
    s_reach[li] = suppress_reach(s_cond, li, len);
  }
  SEL_STAMP(0);
  sel_rank_pass<false>(s_key, s_rank, s_sorted, s_spl, s_m, s_qoff);
  __syncthreads();
  {
    int tie = 0;
#pragma unroll
    for (int j = 0; j < 6; j++)
      for (int t = tid; t < m6[j]; t += SEL_NT) tie |= s_sorted[sp6[j] - a + t] == 0xffffu;
    if (__syncthreads_or(tie)) {  // equal curvatures somewhere in the ring: the index decides (stable sort)
      sel_rank_pass<true>(s_key, s_rank, s_sorted, s_spl, s_m, s_qoff);
      __syncthreads();
    }
  }
  SEL_STAMP(1);
  // dominators: which neighbours come earlier in the point's walk and reach it
  for (int li = S - a + tid; li < E - a; li += SEL_NT) {
    const int j = (li >= s_spl[1]) + (li >= s_spl[2]) + (li >= s_spl[3]) + (li >= s_spl[4]) + (li >= s_spl[5]);
    const int spl = s_spl[j], epl = spl + s_m[j] - 1;
    if (li < spl || li > epl) continue;  // a skipped sector
    const float cv = __uint_as_float(s_key[li]);
    const bool sharp = cv > 0.1, flat = cv < 0.1;
    const int sgn = sharp ? 1 : -1, rk = s_rank[li] * sgn;  // sharp: larger rank first; flat: smaller rank first
    unsigned int dom = 0;
#pragma unroll
    for (int d = 1; d <= 5; d++) {
      const int lo = max(li - d, spl), hi = min(li + d, epl);
      const int rl = s_rank[lo] * sgn, rh = s_rank[hi] * sgn;
      const int xl = s_reach[lo] & 15, xh = s_reach[hi] >> 4;
      const bool bl = (li - d >= spl) & (xl >= d) & (rl > rk);
      const bool bh = (li + d <= epl) & (xh >= d) & (rh > rk);
      dom |= (bl ? 1u : 0u) << (5 - d);
      dom |= (bh ? 1u : 0u) << (5 + d);
    }
    s_dom[li] = (unsigned short)(dom | (sharp ? 1u << 11 : 0u) | (flat ? 1u << 12 : 0u));
  }
  __syncthreads();
  SEL_STAMP(5);
#ifndef LG_SEL_DEBUG
  long long sel_dbg_dummy = 0;
#endif
  {
    int maxm = 0;
#pragma unroll
    for (int j = 0; j < 6; j++) maxm = max(maxm, m6[j]);
    const int nwa = maxm <= SEL_NT ? max((maxm + 31) >> 5, 1) : SEL_NT / 32;  // warps that hold sector points
    if (warp < nwa) {
      if (maxm <= SEL_NT) sel_walks<1>(s_rank, s_dom, s_reach, s_picked, s_label, s_bits, s_sb, s_base, s_spl, s_m, a, my_picks, nwa * 32, &s_phases, SEL_DBG_ARGS);
      else sel_walks<SEL_CAP / SEL_NT>(s_rank, s_dom, s_reach, s_picked, s_label, s_bits, s_sb, s_base, s_spl, s_m, a, my_picks, nwa * 32, &s_phases, SEL_DBG_ARGS);
    }
  }
  __syncthreads();
  const int nsharp = s_base[s_phases][0], nless = s_base[s_phases][1], nflat = s_base[s_phases][2];
  for (int li = tid; li < len; li += SEL_NT) label[a + li] = s_label[li];
  if (any_virtual) {  // what a later, overlapping ring of the reference would find: pick flags and the sorted index order
    for (int li = tid; li < len; li += SEL_NT) picked[a + li] = s_picked[li];
#pragma unroll
    for (int j = 0; j < 6; j++)
      for (int t = tid; t < m6[j]; t += SEL_NT) sort_ind[sp6[j] + t] = a + (int)s_sorted[sp6[j] - a + t];
  }
  // cloudNeighborPicked[0..4] is never re-initialised by the reference (SR:454 starts at 5): marks persist between sweeps
  if (a + tid < 5 && tid < len && s_picked[tid]) stale[a + tid] = 1;
  if (tid == 0) {
    cnt[0] = nsharp;
    cnt[1] = nless;
    cnt[2] = nflat;
    if (too_big) atomicExch(&meta[SRM_ERR], 1);
  }
  SEL_STAMP(5);
#ifdef LG_SEL_DEBUG
  if (blockIdx.x == 1 && blockIdx.y == 0 && tid == 0) {
    for (int k = 0; k < 7; k++) g_sel_dbg[k] += dbg_acc[k];
    g_sel_dbg[7]++;
  }
#endif
}

// Suppression reach with whole-cloud bounds (FENCE (iv) as in oracle/orc_sr.h suppress_neighbours).
__device__ __forceinline__ unsigned char suppress_reach_g(const unsigned char* __restrict__ cond, int i, int n) {
  int nf = 0;
  for (int l = 1; l <= 5; l++) {
    if (i + l >= n) break;
    if (cond[i + l] & C_GAP) break;
    nf = l;
  }
  int nb = 0;
  for (int l = 1; l <= 5; l++) {
    if (i - l < 0) break;
    if (cond[i - l + 1] & C_GAP) break;
    nb = l;
  }
  return (unsigned char)(nf | (nb << 4));
}

constexpr int VIRT_NT = 1024;
constexpr int VIRT_SMEM_KEYS = 16384;  // sectors up to this many points sort in shared memory, longer ones in global scratch
constexpr int VIRT_MAX_KEYS = 32768;

// SR:559-684 for the VIRTUAL rings of a sweep (see sr_select_kernel), replayed with the reference's serial semantics: one
// CTA walks them in ring order on the global state the per-ring CTAs left behind -- cloudSortInd (its current permutation
// decides ties of the stable sort, SR:568-576), cloudNeighborPicked, cloudLabel -- including the five leading entries
// the reference never re-initialises (static arrays SR:68-74: curvature 0, sort index 0, pick flag / label from earlier
// sweeps; `stale`).  Per ring it leaves the picks and the less-flat candidates (label <= 0 at that moment, SR:670-674),
// compacted in index order, for the per-ring voxel grid.
__global__ void __launch_bounds__(VIRT_NT) sr_virtual_kernel(SrParams prm, const float4* __restrict__ c, int* __restrict__ meta,
                                                              const float* __restrict__ curv, const unsigned char* __restrict__ cond,
                                                              unsigned char* picked_g, signed char* label, int* sort_ind,
                                                              unsigned char* __restrict__ reach, unsigned char* stale, int* __restrict__ picks,
                                                              float4* __restrict__ lf_stage, int* __restrict__ lf_meta,
                                                              unsigned long long* gkeys) {
  extern __shared__ unsigned long long vkeys[];
  __shared__ int s_scan[VIRT_NT / 32 + 2];
  volatile unsigned char* picked = picked_g;
  const int tid = threadIdx.x, lane = tid & 31;
  const int n = meta[SRM_N_FULL];
  const int R = prm.n_scans;
  for (int i = tid; i < n; i += VIRT_NT) reach[i] = suppress_reach_g(cond, i, n);
  if (tid < 5 && tid < n) {
    if (stale[tid]) picked[tid] = 1;
    label[tid] = (signed char)stale[8 + tid];
    sort_ind[tid] = 0;
  }
  __syncthreads();
  int lf_off = 0;
  for (int r = 1; r < R; r++) {
    const int S = meta[SRM_SCAN_START + r];
    const int E = (r == R - 1) ? n - 5 : meta[SRM_SCAN_END + r];
    if (tid == 0) lf_meta[2 * r] = lf_meta[2 * r + 1] = 0;
    if (!(S == 0 && E > 0)) continue;
    int* my_picks = picks + r * SR_PICKS_PER_RING;
    int nsharp = 0, nless = 0, nflat = 0;
    int base = lf_off;
    for (int j = 0; j < 6; j++) {
      const int sp = (int)(((long long)E * j) / 6);
      const int ep = (int)(((long long)E * (j + 1)) / 6) - 1;
      const int m = ep - sp + 1;
      if (m <= 0) continue;
      if (m > VIRT_MAX_KEYS) {
        if (tid == 0) atomicExch(&meta[SRM_ERR], 1);
        continue;
      }
      int P = 2;
      while (P < m) P <<= 1;
      unsigned long long* sk = P <= VIRT_SMEM_KEYS ? vkeys : gkeys;
      for (int t = tid; t < P; t += VIRT_NT)
        sk[t] = t < m ? (((unsigned long long)__float_as_uint(curv[sort_ind[sp + t]]) << 32) | (unsigned int)t) : ~0ull;
      __syncthreads();
      for (int k = 2; k <= P; k <<= 1)
        for (int jj = k >> 1; jj > 0; jj >>= 1) {
          for (int t = tid; t < (P >> 1); t += VIRT_NT) {
            const int i = ((t & ~(jj - 1)) << 1) | (t & (jj - 1));
            const int l = i | jj;
            const unsigned long long x = sk[i], y = sk[l];
            const bool asc = (i & k) == 0;
            if ((x > y) == asc) {
              sk[i] = y;
              sk[l] = x;
            }
          }
          __syncthreads();
        }
      {  // position in the previous order -> point index; the sorted order becomes the new cloudSortInd
        int tmp[VIRT_MAX_KEYS / VIRT_NT];
        int q = 0;
        for (int t = tid; t < m; t += VIRT_NT, q++) tmp[q] = sort_ind[sp + (int)(unsigned int)(sk[t] & 0xffffffffull)];
        __syncthreads();
        q = 0;
        for (int t = tid; t < m; t += VIRT_NT, q++) {
          sort_ind[sp + t] = tmp[q];
          sk[t] = (sk[t] & 0xffffffff00000000ull) | (unsigned int)tmp[q];
        }
        __syncthreads();
      }
      if (tid < 32) {
        // ---- SR:578-624: walk down from the largest curvature
        int count = 0;
        bool done = false;
        for (int k = m - 1; k >= 0 && !done; k -= 32) {
          const int kk = k - lane;
          const bool have = kk >= 0;
          const unsigned long long key = have ? sk[kk] : 0ull;
          const float cv = __uint_as_float((unsigned int)(key >> 32));
          const int li = (int)(unsigned int)(key & 0xffffffffull);
          const bool pass = have && (cv > 0.1);
          const unsigned int pm = __ballot_sync(0xffffffffu, pass);
          int nf = 0, nb = 0;
          if (pass) {
            const unsigned char rr = reach[li];
            nf = rr & 15;
            nb = rr >> 4;
          }
          const int lo = li - nb, hi = li + nf;
          const bool alive = pass && picked[li] == 0;
          unsigned int am = __ballot_sync(0xffffffffu, alive);
          int mynum = 0;
          while (am) {
            const int p = __ffs(am) - 1;
            am &= ~(1u << p);
            count++;
            if (count > 20) {  // SR:592-594: the 21st candidate only ends the walk
              done = true;
              break;
            }
            if (lane == p) mynum = count;
            const int plo = __shfl_sync(0xffffffffu, lo, p), phi = __shfl_sync(0xffffffffu, hi, p);
            am &= ~__ballot_sync(0xffffffffu, alive && lane > p && li >= plo && li <= phi);
          }
          if (mynum > 0) {
            if (mynum <= 16) {
              label[li] = 2;
              my_picks[SR_PICK_SHARP + nsharp + mynum - 1] = li;
            } else {
              label[li] = 1;
            }
            my_picks[SR_PICK_LESS + nless + mynum - 1] = li;
            for (int l = -nb; l <= nf; l++) picked[li + l] = 1;
          }
          __syncwarp();
          if (pm != 0xffffffffu) done = true;  // sorted: nothing below the first c <= 0.1 can pass
        }
        const int npick = min(count, 20);
        nsharp += min(npick, 16);
        nless += npick;
        // ---- SR:626-668: walk up from the smallest curvature
        count = 0;
        done = false;
        for (int k = 0; k < m && !done; k += 32) {
          const int kk = k + lane;
          const bool have = kk < m;
          const unsigned long long key = have ? sk[kk] : 0ull;
          const float cv = __uint_as_float((unsigned int)(key >> 32));
          const int li = (int)(unsigned int)(key & 0xffffffffull);
          const bool pass = have && (cv < 0.1);
          const unsigned int pm = __ballot_sync(0xffffffffu, pass);
          int nf = 0, nb = 0;
          if (pass) {
            const unsigned char rr = reach[li];
            nf = rr & 15;
            nb = rr >> 4;
          }
          const int lo = li - nb, hi = li + nf;
          const bool alive = pass && picked[li] == 0;
          unsigned int am = __ballot_sync(0xffffffffu, alive);
          int mynum = 0;
          bool last32 = false;
          while (am) {
            const int p = __ffs(am) - 1;
            am &= ~(1u << p);
            count++;
            if (lane == p) mynum = count;
            if (count >= 32) {  // SR:635-638: the 32nd flat point is kept but neither marked nor suppressing
              if (lane == p) last32 = true;
              done = true;
              break;
            }
            const int plo = __shfl_sync(0xffffffffu, lo, p), phi = __shfl_sync(0xffffffffu, hi, p);
            am &= ~__ballot_sync(0xffffffffu, alive && lane > p && li >= plo && li <= phi);
          }
          if (mynum > 0) {
            label[li] = -1;
            my_picks[SR_PICK_FLAT + nflat + mynum - 1] = li;
            if (!last32)
              for (int l = -nb; l <= nf; l++) picked[li + l] = 1;
          }
          __syncwarp();
          if (pm != 0xffffffffu) done = true;
        }
        nflat += count;
      }
      __syncthreads();
      // SR:670-674 runs per SECTOR, right after its picks: the less-flat candidates of [sp, ep] are the points whose label
      // is <= 0 NOW (a later sector of this ring may still label a point in here: its sort indices were permuted by the
      // earlier, differently cut rings)
      for (int k0 = sp; k0 <= ep; k0 += VIRT_NT) {
        const int k = k0 + tid;
        const int flag = (k <= ep && label[k] <= 0) ? 1 : 0;
        int tot;
        const int ex = block_excl_scan<VIRT_NT>(flag, &tot, s_scan);
        if (flag) lf_stage[base + ex] = c[k];
        base += tot;
      }
    }
    if (tid == 0) {
      int* cnt = meta + SRM_PICK_CNT + r * 3;
      cnt[0] = nsharp;
      cnt[1] = nless;
      cnt[2] = nflat;
    }
    if (tid == 0) {
      lf_meta[2 * r] = lf_off;
      lf_meta[2 * r + 1] = base - lf_off;
    }
    lf_off = base;
    __syncthreads();
  }
  if (tid < 5 && tid < n) {
    stale[tid] = picked[tid];
    stale[8 + tid] = (unsigned char)label[tid];
  }
}

// Compacts picks to feature clouds, derives the less-flat mask (SR:670-674) and the per-ring voxel jobs (SR:677-683).
__device__ __forceinline__ void sr_collect_kernel_body(SrParams prm, const float4* __restrict__ c, int* __restrict__ meta,
                                                          const signed char* __restrict__ label, const int* __restrict__ picks,
                                                          float4* __restrict__ sharp, float4* __restrict__ less_sharp, float4* __restrict__ flat,
                                                          unsigned char* __restrict__ lf_valid, float4* __restrict__ lf_tmp, VoxSegD* __restrict__ segs,
                                                          int features_only) {
  const int r = blockIdx.x, tid = threadIdx.x;
  const int n = meta[SRM_N_FULL];
  const int R = prm.n_scans;
  const int S = (r == 0) ? 5 : meta[SRM_SCAN_START + r];
  const int E = (r == R - 1) ? n - 5 : meta[SRM_SCAN_END + r];
  int off[3] = {0, 0, 0}, tot[3] = {0, 0, 0};
  for (int k = 0; k < R; k++)
    for (int t = 0; t < 3; t++) {
      int v = meta[SRM_PICK_CNT + k * 3 + t];
      if (k < r) off[t] += v;
      tot[t] += v;
    }
  const int* my = picks + r * SR_PICKS_PER_RING;
  const int* cnt = meta + SRM_PICK_CNT + r * 3;
  for (int i = tid; i < cnt[0]; i += blockDim.x) sharp[off[0] + i] = c[my[SR_PICK_SHARP + i]];
  for (int i = tid; i < cnt[1]; i += blockDim.x) less_sharp[off[1] + i] = c[my[SR_PICK_LESS + i]];
  for (int i = tid; i < cnt[2]; i += blockDim.x) flat[off[2] + i] = c[my[SR_PICK_FLAT + i]];
  if (features_only) {  // second pass after sr_virtual_kernel: the voxel jobs of the first pass stand
    if (tid == 0 && r == 0) {
      meta[SRM_N_SHARP] = tot[0];
      meta[SRM_N_LESS_SHARP] = tot[1];
      meta[SRM_N_FLAT] = tot[2];
    }
    return;
  }
  const bool is_virtual = r > 0 && S == 0 && E > 0;  // handled by sr_virtual_kernel (empty job here)
  const int a = is_virtual ? 0 : max(S - 5, 0), b = is_virtual ? 0 : min(E + 5, n);
  for (int i = a + tid; i < b; i += blockDim.x) lf_valid[i] = (i >= S && i < E && label[i] <= 0) ? 1 : 0;
  if (tid == 0) {
    VoxSegD sg;
    sg.in = c + a;
    sg.valid = lf_valid + a;
    sg.out = lf_tmp + a;
    sg.out_count = meta + SRM_LF_CNT + r;
    sg.n = max(b - a, 0);
    sg.leaf = 0.2f;
    segs[r] = sg;
    if (r == 0) {
      meta[SRM_N_SHARP] = tot[0];
      meta[SRM_N_LESS_SHARP] = tot[1];
      meta[SRM_N_FLAT] = tot[2];
    }
    if (r == 0) meta[SRM_SCAN_START + 0] = 5;        // SR:489 (diagnostics see the final values)
    if (r == R - 1) meta[SRM_SCAN_END + R - 1] = n - 5;  // SR:490
  }
}

__device__ __forceinline__ void sr_concat_kernel_body(SrParams prm, int* __restrict__ meta, const VoxSegD* __restrict__ segs,
                                                         float4* __restrict__ less_flat) {
  const int r = blockIdx.x, tid = threadIdx.x;
  const int R = prm.n_scans;
  int off = 0, tot = 0;
  for (int k = 0; k < R; k++) {
    int v = max(meta[SRM_LF_CNT + k], 0);
    if (k < r) off += v;
    tot += v;
  }
  const int cnt = max(meta[SRM_LF_CNT + r], 0);
  const float4* src = segs[r].out;
  for (int i = tid; i < cnt; i += blockDim.x) less_flat[off + i] = src[i];
  if (r == 0 && tid == 0) {
    meta[SRM_N_LESS_FLAT] = tot;
    meta[SRM_JSTAR] = 0x7fffffff;  // re-arm the prefix-min for the next sweep
  }
}

// ---- SR:364-434, the IMU branch: every kept point advances imuPointerFront, interpolates the "Cur" IMU values at its own
// time, and is rotated / shifted back to the sweep's start (SR:121-184).  The reference does this point by point with
// state that persists from point to point; here ONE CTA walks the sweep in chunks of 1024 points in input order:
//   imuPointerFront is monotone: front_i = max(front_(i-1), g_i), g_i = first ring offset whose stamp exceeds the point's
//     time (exact while the ring's stamps increase with the offset, as IMU streams do) -> an inclusive max-scan;
//   the Cur values are those of the LAST point whose update passed its 0.2 s gate (SR:375, 390) -> a second max-scan over
//     "index if valid"; a point then recomputes that source point's interpolation (same arithmetic, same result);
//   the Start values are latched at the sweep's first finite point if it is kept (SR:410-421), which precedes every
//     other kept point; ShiftToStartIMU / VeloToStartIMU / TransformToStartIMU are pure functions of (Cur_i, Start, time_i).
// sinf / cosf are the libm-exact ports (lg_libm.cuh), products are not contracted: bit-identical to the host arithmetic.
constexpr int IMU_NT = 1024;
struct ImuCur { float v[9]; };
__device__ __forceinline__ int imu_scan_max(int v, int carry, int* s_w /* >= 33 */) {  // inclusive max-scan over the CTA + carry-in; ends with a barrier
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v = max(v, t);
  }
  if (lane == 31) s_w[w] = v;
  __syncthreads();
  if (w == 0) {
    int x = s_w[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x = max(x, t);
    }
    s_w[lane] = x;
  }
  __syncthreads();
  int r = max(v, carry);
  if (w > 0) r = max(r, s_w[w - 1]);
  __syncthreads();
  return r;
}
// the interpolation of SR:373-408 for a point at time t whose imuPointerFront is ring index F; false = its gate failed
__device__ __forceinline__ bool imu_interp(const SrImuRing& R, double t, int F, ImuCur& c) {
  const double tf = R.time[F];
  if (t > tf) {
    if (!(t - tf < 0.2)) return false;
    c.v[0] = R.roll[F]; c.v[1] = R.pitch[F]; c.v[2] = R.yaw[F];
    c.v[3] = R.veloX[F]; c.v[4] = R.veloY[F]; c.v[5] = R.veloZ[F];
    c.v[6] = R.shiftX[F]; c.v[7] = R.shiftY[F]; c.v[8] = R.shiftZ[F];
    return true;
  }
  if (!(tf - t < 0.2)) return false;
  const int B = (F + SR_IMU_Q - 1) % SR_IMU_Q;
  const float rf = (float)((t - R.time[B]) / (tf - R.time[B]));
  const float rb = (float)((tf - t) / (tf - R.time[B]));
  c.v[0] = R.roll[F] * rf + R.roll[B] * rb;
  c.v[1] = R.pitch[F] * rf + R.pitch[B] * rb;
  if (R.yaw[F] - R.yaw[B] > M_PI) {
    c.v[2] = (float)(R.yaw[F] * rf + (R.yaw[B] + 2 * M_PI) * rb);
  } else if (R.yaw[F] - R.yaw[B] < -M_PI) {
    c.v[2] = (float)(R.yaw[F] * rf + (R.yaw[B] - 2 * M_PI) * rb);
  } else {
    c.v[2] = R.yaw[F] * rf + R.yaw[B] * rb;
  }
  c.v[3] = R.veloX[F] * rf + R.veloX[B] * rb;
  c.v[4] = R.veloY[F] * rf + R.veloY[B] * rb;
  c.v[5] = R.veloZ[F] * rf + R.veloZ[B] * rb;
  c.v[6] = R.shiftX[F] * rf + R.shiftX[B] * rb;
  c.v[7] = R.shiftY[F] * rf + R.shiftY[B] * rb;
  c.v[8] = R.shiftZ[F] * rf + R.shiftZ[B] * rb;
  return true;
}
// SR:121-160: the shift / velocity of the point relative to the sweep's start, in the start frame
__device__ __forceinline__ void imu_from_start(const ImuCur& c, const float* S, float pointTime, float* sh, float* ve) {
  const float cy = lgm_cosf(S[2]), sy = lgm_sinf(S[2]), cp = lgm_cosf(S[1]), sp = lgm_sinf(S[1]), cr = lgm_cosf(S[0]), sr = lgm_sinf(S[0]);
  {
    const float fx = c.v[6] - S[6] - S[3] * pointTime, fy = c.v[7] - S[7] - S[4] * pointTime, fz = c.v[8] - S[8] - S[5] * pointTime;
    const float x1 = cy * fx - sy * fz, y1 = fy, z1 = sy * fx + cy * fz;
    const float x2 = x1, y2 = cp * y1 + sp * z1, z2 = -sp * y1 + cp * z1;
    sh[0] = cr * x2 + sr * y2;
    sh[1] = -sr * x2 + cr * y2;
    sh[2] = z2;
  }
  {
    const float fx = c.v[3] - S[3], fy = c.v[4] - S[4], fz = c.v[5] - S[5];
    const float x1 = cy * fx - sy * fz, y1 = fy, z1 = sy * fx + cy * fz;
    const float x2 = x1, y2 = cp * y1 + sp * z1, z2 = -sp * y1 + cp * z1;
    ve[0] = cr * x2 + sr * y2;
    ve[1] = -sr * x2 + cr * y2;
    ve[2] = z2;
  }
}
__global__ void __launch_bounds__(IMU_NT) sr_imu_kernel(SrParams prm, const float* __restrict__ xyz, int n, int stride_bytes,
                                                         const signed char* __restrict__ ring8, const float* __restrict__ ori_raw,
                                                         const int* __restrict__ meta, SrImuJob J, float* __restrict__ pts,
                                                         double* __restrict__ g_t, int2* __restrict__ g_fs) {
  __shared__ SrImuRing R;
  __shared__ int s_w[33];
  __shared__ float s_ori[2];
  __shared__ int s_ok, s_f0;
  __shared__ float s_start[9];
  const int tid = threadIdx.x;
  for (int k = tid; k < (int)(sizeof(SrImuRing) / 4); k += IMU_NT) reinterpret_cast<int*>(&R)[k] = reinterpret_cast<const int*>(J.ring)[k];
  if (tid == 0) {
    sweep_ori(xyz, n, stride_bytes, &s_ori[0], &s_ori[1], &s_ok);
    s_f0 = 0x7fffffff;
  }
  if (tid < 9) s_start[tid] = J.carry->start[tid];
  __syncthreads();
  {  // i == 0 of the NaN-filtered cloud (SR:410): the first finite point
    int f = 0x7fffffff;
    for (int i = tid; i < n && f == 0x7fffffff; i += IMU_NT)
      if (finite3(pt_at(xyz, stride_bytes, i))) f = i;
    if (f != 0x7fffffff) atomicMin(&s_f0, f);
  }
  __syncthreads();
  const int f0 = s_f0;
  const float startOri = s_ori[0], endOri = s_ori[1];
  const int jstar = meta[SRM_JSTAR];
  const int front0 = J.carry->front;
  const int L = (J.last - front0 + SR_IMU_Q) % SR_IMU_Q;  // offsets 0 .. L: the ring entries the pointer can reach
  auto point_time = [&](int i) {  // relTime as sr_scatter_kernel forms it (SR:341-362), times scanPeriod (SR:365)
    float ori = ori_raw[i];
    if (i <= jstar) {
      if (ori < startOri - M_PI / 2) {
        ori = (float)(ori + 2 * M_PI);
      } else if (ori > startOri + M_PI * 3 / 2) {
        ori = (float)(ori - 2 * M_PI);
      }
    } else {
      ori = (float)(ori + 2 * M_PI);
      if (ori < endOri - M_PI * 3 / 2) {
        ori = (float)(ori + 2 * M_PI);
      } else if (ori > endOri + M_PI / 2) {
        ori = (float)(ori - 2 * M_PI);
      }
    }
    const float relTime = (ori - startOri) / (endOri - startOri);
    return (float)(relTime * prm.scan_period);
  };
  ImuCur carry_cur;
#pragma unroll
  for (int k = 0; k < 9; k++) carry_cur.v[k] = J.carry->cur[k];
  int cF = 0, cS = -1, cK = -1;  // carries: front offset, last valid source, last kept point
  for (int base = 0; base < n; base += IMU_NT) {
    const int i = base + tid;
    const bool kept = s_ok && i < n && ring8[i] >= 0;
    float pointTime = 0.f;
    double t = 0.0;
    int g = 0;
    if (kept) {
      pointTime = point_time(i);
      t = J.time_scan + pointTime;
      while (g < L && !(t < R.time[(front0 + g) % SR_IMU_Q])) g++;  // SR:366-371 from the sweep's first pointer
    }
    const int F = imu_scan_max(g, cF, s_w);
    ImuCur mine;
    bool valid = false;
    if (kept) valid = imu_interp(R, t, (front0 + F) % SR_IMU_Q, mine);
    const int S = imu_scan_max(valid ? i : -1, cS, s_w);
    const int K = imu_scan_max(kept ? i : -1, cK, s_w);
    if (i < n) {
      g_t[i] = t;
      g_fs[i] = make_int2(F, S);
    }
    __syncthreads();  // the source point of a lane may sit anywhere in this or an earlier chunk
    ImuCur cur = carry_cur;
    if (kept) {
      if (S == i) cur = mine;
      else if (S >= 0) imu_interp(R, g_t[S], (front0 + g_fs[S].x) % SR_IMU_Q, cur);
      if (i == f0) {
#pragma unroll
        for (int k = 0; k < 9; k++) s_start[k] = cur.v[k];  // SR:410-421
      }
    }
    __syncthreads();
    if (kept) {
      const float* p = pt_at(xyz, stride_bytes, i);
      float px = p[1], py = p[2], pz = p[0];  // SR:293-295
      if (i != f0) {
        float sh[3], ve[3];
        imu_from_start(cur, s_start, pointTime, sh, ve);
        // TransformToStartIMU SR:163-184
        const float crc = lgm_cosf(cur.v[0]), src = lgm_sinf(cur.v[0]), cpc = lgm_cosf(cur.v[1]), spc = lgm_sinf(cur.v[1]);
        const float cyc = lgm_cosf(cur.v[2]), syc = lgm_sinf(cur.v[2]);
        const float cys = lgm_cosf(s_start[2]), sys = lgm_sinf(s_start[2]), cps = lgm_cosf(s_start[1]), sps = lgm_sinf(s_start[1]);
        const float crs = lgm_cosf(s_start[0]), srs = lgm_sinf(s_start[0]);
        const float x1 = crc * px - src * py, y1 = src * px + crc * py, z1 = pz;
        const float x2 = x1, y2 = cpc * y1 - spc * z1, z2 = spc * y1 + cpc * z1;
        const float x3 = cyc * x2 + syc * z2, y3 = y2, z3 = -syc * x2 + cyc * z2;
        const float x4 = cys * x3 - sys * z3, y4 = y3, z4 = sys * x3 + cys * z3;
        const float x5 = x4, y5 = cps * y4 + sps * z4, z5 = -sps * y4 + cps * z4;
        px = crs * x5 + srs * y5 + sh[0];
        py = -srs * x5 + crs * y5 + sh[1];
        pz = z5 + sh[2];
      }
      pts[3 * (size_t)i] = px;
      pts[3 * (size_t)i + 1] = py;
      pts[3 * (size_t)i + 2] = pz;
    }
    // carries of this chunk (its last lane holds the inclusive results)
    __shared__ int s_c[3];
    if (tid == IMU_NT - 1) {
      s_c[0] = F; s_c[1] = S; s_c[2] = K;
    }
    __syncthreads();
    cF = s_c[0]; cS = s_c[1]; cK = s_c[2];
    __syncthreads();
  }
  // ---- what persists into the next sweep (and goes out as /imu_trans, SR:730-745): the state after the LAST kept point
  if (tid == 0) {
    SrImuCarry out = *J.carry;
    if (cK >= 0) {
      ImuCur cur = carry_cur;
      const int S = g_fs[cK].y;
      if (S >= 0) imu_interp(R, g_t[S], (front0 + g_fs[S].x) % SR_IMU_Q, cur);
#pragma unroll
      for (int k = 0; k < 9; k++) {
        out.cur[k] = cur.v[k];
        out.start[k] = s_start[k];
      }
      if (cK != f0) {
        const float pointTime = point_time(cK);
        imu_from_start(cur, s_start, pointTime, out.shift_from_start, out.velo_from_start);
      }
      out.front = (front0 + cF) % SR_IMU_Q;
    }
    *J.carry = out;
  }
}

// ------------------------------------------------------------------------------------------------ launch forms
// Every kernel above exists twice: for ONE sequence (arguments by value) and BATCHED over several sequences -- grid.y is
// the sequence, whose arguments come from a device table (SURVEY 8b `*_batch`): eight launches extract B sweeps instead
// of 8 B.  grid.x is the largest any member needs; a CTA beyond its member's own grid returns at once.
struct SrK {
  SrParams prm;
  const float* xyz;
  int n, stride, nblocks;
  signed char* ring8;
  float* ori_raw;
  unsigned int* hist;
  int* meta;
  float4* full;
  float* curv;
  unsigned char *cond, *picked, *mask_diag;
  signed char* label;
  int *picks, *sort_ind;
  unsigned char* stale;
  float4 *sharp, *less_sharp, *flat;
  unsigned char* lf_valid;
  float4* lf_tmp;
  VoxSegD* segs;
  float4* less_flat;
};

__global__ void __launch_bounds__(SR_NT) sr_ring_kernel(SrParams prm, const float* __restrict__ xyz, int n, int stride_bytes,
                                                         signed char* __restrict__ ring8, float* __restrict__ ori_raw,
                                                         unsigned int* __restrict__ hist, int nblocks, int* __restrict__ meta) {
  sr_ring_kernel_body(prm, xyz, n, stride_bytes, ring8, ori_raw, hist, nblocks, meta);
}
__global__ void __launch_bounds__(SR_NT) sr_ring_batch_kernel(const SrK* __restrict__ tab) {
  const SrK A = tab[blockIdx.y];
  if ((int)blockIdx.x >= A.nblocks) return;
  sr_ring_kernel_body(A.prm, A.xyz, A.n, A.stride, A.ring8, A.ori_raw, A.hist, A.nblocks, A.meta);
}
__global__ void __launch_bounds__(1024) sr_scan_kernel(SrParams prm, unsigned int* __restrict__ hist, int nblocks, int* __restrict__ meta) {
  sr_scan_kernel_body(prm, hist, nblocks, meta);
}
__global__ void __launch_bounds__(1024) sr_scan_batch_kernel(const SrK* __restrict__ tab) {
  const SrK A = tab[blockIdx.y];
  sr_scan_kernel_body(A.prm, A.hist, A.nblocks, A.meta);
}
__global__ void __launch_bounds__(SR_NT) sr_scatter_kernel(SrParams prm, const float* __restrict__ xyz, int n, int stride_bytes,
                                                            const signed char* __restrict__ ring8, const float* __restrict__ ori_raw,
                                                            const unsigned int* __restrict__ hist, int nblocks, const int* __restrict__ meta,
                                                            float4* __restrict__ full, const float* __restrict__ imu_pts) {
  sr_scatter_kernel_body(prm, xyz, n, stride_bytes, ring8, ori_raw, hist, nblocks, meta, full, imu_pts);
}
__global__ void __launch_bounds__(SR_NT) sr_scatter_batch_kernel(const SrK* __restrict__ tab) {
  const SrK A = tab[blockIdx.y];
  if ((int)blockIdx.x >= A.nblocks) return;
  sr_scatter_kernel_body(A.prm, A.xyz, A.n, A.stride, A.ring8, A.ori_raw, A.hist, A.nblocks, A.meta, A.full);
}
__global__ void __launch_bounds__(256) sr_curv_kernel(SrParams prm, const float4* __restrict__ c, int* __restrict__ meta,
                                                       float* __restrict__ curv, unsigned char* __restrict__ cond, signed char* __restrict__ label) {
  sr_curv_kernel_body(prm, c, meta, curv, cond, label);
}
__global__ void __launch_bounds__(256) sr_curv_batch_kernel(const SrK* __restrict__ tab) {
  const SrK A = tab[blockIdx.y];
  if ((int)(blockIdx.x * blockDim.x) >= A.n) return;
  sr_curv_kernel_body(A.prm, A.full, A.meta, A.curv, A.cond, A.label);
}
__global__ void __launch_bounds__(SEL_NT, 2) sr_select_kernel(SrParams prm, const float4* __restrict__ c, int* __restrict__ meta,
                                                            const float* __restrict__ curv, const unsigned char* __restrict__ cond,
                                                            unsigned char* __restrict__ picked, unsigned char* __restrict__ mask_diag,
                                                            signed char* __restrict__ label, int* __restrict__ picks,
                                                            int* __restrict__ sort_ind, unsigned char* __restrict__ stale) {
  sr_select_kernel_body(prm, c, meta, curv, cond, picked, mask_diag, label, picks, sort_ind, stale);
}
__global__ void __launch_bounds__(SEL_NT, 2) sr_select_batch_kernel(const SrK* __restrict__ tab) {
  const SrK A = tab[blockIdx.y];
  if ((int)blockIdx.x >= A.prm.n_scans) return;
  sr_select_kernel_body(A.prm, A.full, A.meta, A.curv, A.cond, A.picked, A.mask_diag, A.label, A.picks, A.sort_ind, A.stale);
}
__global__ void __launch_bounds__(256) sr_collect_kernel(SrParams prm, const float4* __restrict__ c, int* __restrict__ meta,
                                                          const signed char* __restrict__ label, const int* __restrict__ picks,
                                                          float4* __restrict__ sharp, float4* __restrict__ less_sharp, float4* __restrict__ flat,
                                                          unsigned char* __restrict__ lf_valid, float4* __restrict__ lf_tmp, VoxSegD* __restrict__ segs,
                                                          int features_only) {
  sr_collect_kernel_body(prm, c, meta, label, picks, sharp, less_sharp, flat, lf_valid, lf_tmp, segs, features_only);
}
__global__ void __launch_bounds__(256) sr_collect_batch_kernel(const SrK* __restrict__ tab) {
  const SrK A = tab[blockIdx.y];
  if ((int)blockIdx.x >= A.prm.n_scans) return;
  sr_collect_kernel_body(A.prm, A.full, A.meta, A.label, A.picks, A.sharp, A.less_sharp, A.flat, A.lf_valid, A.lf_tmp, A.segs, 0);
}
__global__ void __launch_bounds__(256) sr_concat_kernel(SrParams prm, int* __restrict__ meta, const VoxSegD* __restrict__ segs,
                                                         float4* __restrict__ less_flat) {
  sr_concat_kernel_body(prm, meta, segs, less_flat);
}
__global__ void __launch_bounds__(256) sr_concat_batch_kernel(const SrK* __restrict__ tab) {
  const SrK A = tab[blockIdx.y];
  if ((int)blockIdx.x >= A.prm.n_scans) return;
  sr_concat_kernel_body(A.prm, A.meta, A.segs, A.less_flat);
}

}  // namespace

// buffers of one sequence for a sweep of n points (grown on demand; nothing is allocated in steady state)
static int lg_extract_prepare(SrWs& ws, const SrParams& prm, int n, cudaStream_t st, int* nblocks_out) {
  const int R = prm.n_scans;
  if (R > MAXR || R < 1) return LOAM_EINVAL;
  const int nblocks = std::max(1, lg_div_up(n, SR_NT * sr_items_for(n)));
  *nblocks_out = nblocks;
  LG_CHECK(ws.ring8.ensure((size_t)n + 16, st));
  LG_CHECK(ws.ori_raw.ensure((size_t)(n + 16) * 4, st));
  LG_CHECK(ws.hist.ensure((size_t)R * nblocks * 4, st));
  LG_CHECK(ws.full.ensure((size_t)(n + 16) * 16, st));
  LG_CHECK(ws.curv.ensure((size_t)(n + 16) * 4, st));
  LG_CHECK(ws.cond.ensure((size_t)n + 16, st));
  LG_CHECK(ws.picked.ensure((size_t)n + 16, st));
  LG_CHECK(ws.mask_diag.ensure((size_t)n + 16, st));
  LG_CHECK(ws.label.ensure((size_t)n + 16, st));
  LG_CHECK(ws.lf_valid.ensure((size_t)n + 16, st));
  LG_CHECK(ws.lf_tmp.ensure((size_t)(n + 16) * 16, st));
  LG_CHECK(ws.less_flat.ensure((size_t)(n + 16) * 16, st));
  LG_CHECK(ws.picks.ensure((size_t)R * SR_PICKS_PER_RING * 4, st));
  LG_CHECK(ws.sharp.ensure((size_t)R * 96 * 16, st));
  LG_CHECK(ws.less_sharp.ensure((size_t)R * 120 * 16, st));
  LG_CHECK(ws.flat.ensure((size_t)R * 192 * 16, st));
  LG_CHECK(ws.segs.ensure((size_t)R * sizeof(VoxSegD), st));
  LG_CHECK(ws.sort_ind.ensure((size_t)(n + 16) * 4, st));
  if (!ws.stale.p) {
    LG_CHECK(ws.stale.ensure(16, st));
    LG_CHECK(cudaMemsetAsync(ws.stale.p, 0, 16, st));
  }
  if (!ws.meta.p) {
    LG_CHECK(ws.meta.ensure(SRM_SIZE * 4, st));
    LG_CHECK(cudaMemsetAsync(ws.meta.p, 0, SRM_SIZE * 4, st));
    int big = 0x7fffffff;
    LG_CHECK(cudaMemcpyAsync(ws.meta.as<int>() + SRM_JSTAR, &big, 4, cudaMemcpyHostToDevice, st));
    LG_CHECK(cudaStreamSynchronize(st));
  }
  return LOAM_OK;
}

int lg_extract_launch(SrWs& ws, const SrParams& prm, const float* d_xyz, int n, int stride_bytes, cudaStream_t st, long long* launches,
                      const SrImuJob* imu) {
  const int R = prm.n_scans;
  int nblocks = 1;
  {
    int rcp = lg_extract_prepare(ws, prm, n, st, &nblocks);
    if (rcp) return rcp;
  }
  int* meta = ws.meta.as<int>();
  if (n <= 0) {
    LG_CHECK(cudaMemsetAsync(meta + SRM_N_FULL, 0, 5 * 4, st));
    return LOAM_OK;
  }
  {
  LgProfScope prof_scope(LGK_EXTRACT, st, (double)n);
  sr_ring_kernel<<<nblocks, SR_NT, 0, st>>>(prm, d_xyz, n, stride_bytes, ws.ring8.as<signed char>(), ws.ori_raw.as<float>(),
                                            ws.hist.as<unsigned int>(), nblocks, meta);
  sr_scan_kernel<<<1, 1024, 0, st>>>(prm, ws.hist.as<unsigned int>(), nblocks, meta);
  const float* imu_pts = nullptr;
  if (imu) {  // SR:364-434: de-skew the kept points (input order) before they are bucketed
    LG_CHECK(ws.imu_pts.ensure((size_t)(n + 16) * 12, st));
    LG_CHECK(ws.imu_t.ensure((size_t)(n + 16) * 8, st));
    LG_CHECK(ws.imu_fs.ensure((size_t)(n + 16) * 8, st));
    sr_imu_kernel<<<1, IMU_NT, 0, st>>>(prm, d_xyz, n, stride_bytes, ws.ring8.as<signed char>(), ws.ori_raw.as<float>(), meta, *imu,
                                        ws.imu_pts.as<float>(), ws.imu_t.as<double>(), ws.imu_fs.as<int2>());
    (*launches)++;
    imu_pts = ws.imu_pts.as<float>();
  }
  sr_scatter_kernel<<<nblocks, SR_NT, 0, st>>>(prm, d_xyz, n, stride_bytes, ws.ring8.as<signed char>(), ws.ori_raw.as<float>(),
                                               ws.hist.as<unsigned int>(), nblocks, meta, ws.full.as<float4>(), imu_pts);
  sr_curv_kernel<<<lg_div_up(n, 256), 256, 0, st>>>(prm, ws.full.as<float4>(), meta, ws.curv.as<float>(), ws.cond.as<unsigned char>(),
                                                    ws.label.as<signed char>());
  }
  {
  LgProfScope prof_scope(LGK_SR_SELECT, st, (double)n);
  static bool sel_attr[64] = {};  // the opt-in is per device
  constexpr int sel_smem = SEL_SMEM;
  int dev = 0;
  LG_CHECK(cudaGetDevice(&dev));
  if (!sel_attr[dev & 63]) {
    LG_CHECK(cudaFuncSetAttribute(sr_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, sel_smem));
    sel_attr[dev & 63] = true;
  }
  sr_select_kernel<<<R, SEL_NT, sel_smem, st>>>(prm, ws.full.as<float4>(), meta, ws.curv.as<float>(), ws.cond.as<unsigned char>(),
                                      ws.picked.as<unsigned char>(), ws.mask_diag.as<unsigned char>(), ws.label.as<signed char>(),
                                      ws.picks.as<int>(), ws.sort_ind.as<int>(), ws.stale.as<unsigned char>());
  }
  LgProfScope prof_scope(LGK_EXTRACT, st, 0.0);
  sr_collect_kernel<<<R, 256, 0, st>>>(prm, ws.full.as<float4>(), meta, ws.label.as<signed char>(), ws.picks.as<int>(),
                                       ws.sharp.as<float4>(), ws.less_sharp.as<float4>(), ws.flat.as<float4>(),
                                       ws.lf_valid.as<unsigned char>(), ws.lf_tmp.as<float4>(), ws.segs.as<VoxSegD>(), 0);
  (*launches) += 6;
  int rc = lg_vox_small(ws.segs.as<VoxSegD>(), R, (n / R) * 2 + 64 <= 4096 ? 4096 : 16384, meta + SRM_VOX_OVERFLOW, st, launches);
  if (rc) return rc;
  sr_concat_kernel<<<R, 256, 0, st>>>(prm, meta, ws.segs.as<VoxSegD>(), ws.less_flat.as<float4>());
  (*launches)++;
  LG_CHECK(cudaGetLastError());
  return LOAM_OK;
}

int lg_extract_virtual_launch(SrWs& ws, const SrParams& prm, int n, size_t stage_points, cudaStream_t st, long long* launches) {
  const int R = prm.n_scans;
  LG_CHECK(ws.reach.ensure((size_t)n + 16, st));
  LG_CHECK(ws.lf_stage.ensure((stage_points + 16) * 16, st));
  LG_CHECK(ws.lf_vout.ensure((stage_points + 16) * 16, st));
  LG_CHECK(ws.lf_meta.ensure((size_t)2 * MAXR * 4, st));
  LG_CHECK(ws.gkeys.ensure((size_t)VIRT_MAX_KEYS * 8, st));
  LG_CHECK(ws.less_flat.ensure(((size_t)n + stage_points + 16) * 16, st));  // virtual rings contribute the rings before them AGAIN
  static bool virt_attr[64] = {};
  constexpr int virt_smem = VIRT_SMEM_KEYS * 8;
  int dev = 0;
  LG_CHECK(cudaGetDevice(&dev));
  if (!virt_attr[dev & 63]) {
    LG_CHECK(cudaFuncSetAttribute(sr_virtual_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, virt_smem));
    virt_attr[dev & 63] = true;
  }
  LgProfScope prof_scope(LGK_SR_SELECT, st, 0.0);
  sr_virtual_kernel<<<1, VIRT_NT, virt_smem, st>>>(prm, ws.full.as<float4>(), ws.meta.as<int>(), ws.curv.as<float>(), ws.cond.as<unsigned char>(),
                                                   ws.picked.as<unsigned char>(), ws.label.as<signed char>(), ws.sort_ind.as<int>(),
                                                   ws.reach.as<unsigned char>(), ws.stale.as<unsigned char>(), ws.picks.as<int>(),
                                                   ws.lf_stage.as<float4>(), ws.lf_meta.as<int>(), ws.gkeys.as<unsigned long long>());
  (*launches)++;
  LG_CHECK(cudaGetLastError());
  (void)R;
  return LOAM_OK;
}

int lg_extract_finish_launch(SrWs& ws, const SrParams& prm, cudaStream_t st, long long* launches) {
  const int R = prm.n_scans;
  LgProfScope prof_scope(LGK_EXTRACT, st, 0.0);
  sr_collect_kernel<<<R, 256, 0, st>>>(prm, ws.full.as<float4>(), ws.meta.as<int>(), ws.label.as<signed char>(), ws.picks.as<int>(),
                                       ws.sharp.as<float4>(), ws.less_sharp.as<float4>(), ws.flat.as<float4>(),
                                       ws.lf_valid.as<unsigned char>(), ws.lf_tmp.as<float4>(), ws.segs.as<VoxSegD>(), 1);
  sr_concat_kernel<<<R, 256, 0, st>>>(prm, ws.meta.as<int>(), ws.segs.as<VoxSegD>(), ws.less_flat.as<float4>());
  (*launches) += 2;
  LG_CHECK(cudaGetLastError());
  return LOAM_OK;
}

// One sweep of each of B sequences with eight launches in all (SURVEY 8b `*_batch`): see SrK above.  Members must have
// n > 0 (the caller handles empty sweeps); the table travels through `tab` (device) on `st`.
int lg_extract_launch_batch(SrWs* const* ws, const SrParams* prm, const float* const* d_xyz, const int* n, const int* stride_bytes, int B,
                            DevBuf& tab, cudaStream_t st, long long* launches) {
  if (B <= 0) return LOAM_OK;
  std::vector<SrK> host(B);
  int max_blocks = 1, max_R = 1, max_n = 1;
  bool small_rings = true;
  for (int b = 0; b < B; b++) {
    if (n[b] <= 0) return LOAM_EINVAL;
    int nblocks = 1;
    int rc = lg_extract_prepare(*ws[b], prm[b], n[b], st, &nblocks);
    if (rc) return rc;
    SrWs& w = *ws[b];
    SrK& k = host[b];
    k.prm = prm[b];
    k.xyz = d_xyz[b]; k.n = n[b]; k.stride = stride_bytes[b]; k.nblocks = nblocks;
    k.ring8 = w.ring8.as<signed char>(); k.ori_raw = w.ori_raw.as<float>(); k.hist = w.hist.as<unsigned int>(); k.meta = w.meta.as<int>();
    k.full = w.full.as<float4>(); k.curv = w.curv.as<float>(); k.cond = w.cond.as<unsigned char>(); k.picked = w.picked.as<unsigned char>();
    k.mask_diag = w.mask_diag.as<unsigned char>(); k.label = w.label.as<signed char>(); k.picks = w.picks.as<int>();
    k.sort_ind = w.sort_ind.as<int>(); k.stale = w.stale.as<unsigned char>(); k.sharp = w.sharp.as<float4>();
    k.less_sharp = w.less_sharp.as<float4>(); k.flat = w.flat.as<float4>(); k.lf_valid = w.lf_valid.as<unsigned char>();
    k.lf_tmp = w.lf_tmp.as<float4>(); k.segs = w.segs.as<VoxSegD>(); k.less_flat = w.less_flat.as<float4>();
    max_blocks = std::max(max_blocks, nblocks);
    max_R = std::max(max_R, prm[b].n_scans);
    max_n = std::max(max_n, n[b]);
    small_rings = small_rings && (n[b] / prm[b].n_scans) * 2 + 64 <= 4096;
  }
  LG_CHECK(tab.ensure((size_t)B * sizeof(SrK) + 2 * (size_t)B * sizeof(void*) + 64, st));
  LG_CHECK(cudaMemcpyAsync(tab.p, host.data(), (size_t)B * sizeof(SrK), cudaMemcpyHostToDevice, st));
  const SrK* d_tab = tab.as<SrK>();
  static bool sel_attr[64] = {};
  constexpr int sel_smem = SEL_SMEM;
  int dev = 0;
  LG_CHECK(cudaGetDevice(&dev));
  if (!sel_attr[dev & 63]) {
    LG_CHECK(cudaFuncSetAttribute(sr_select_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, sel_smem));
    sel_attr[dev & 63] = true;
  }
  double units = 0;
  for (int b = 0; b < B; b++) units += n[b];
  {
    LgProfScope prof_scope(LGK_EXTRACT, st, units);
    sr_ring_batch_kernel<<<dim3(max_blocks, B), SR_NT, 0, st>>>(d_tab);
    sr_scan_batch_kernel<<<dim3(1, B), 1024, 0, st>>>(d_tab);
    sr_scatter_batch_kernel<<<dim3(max_blocks, B), SR_NT, 0, st>>>(d_tab);
    sr_curv_batch_kernel<<<dim3(lg_div_up(max_n, 256), B), 256, 0, st>>>(d_tab);
  }
  {
    LgProfScope prof_scope(LGK_SR_SELECT, st, units);
    sr_select_batch_kernel<<<dim3(max_R, B), SEL_NT, sel_smem, st>>>(d_tab);
  }
  LgProfScope prof_scope(LGK_EXTRACT, st, 0.0);
  sr_collect_batch_kernel<<<dim3(max_R, B), 256, 0, st>>>(d_tab);
  (*launches) += 6;
  // per-ring voxel grids of all members in one launch: a table of {segment array, overflow flag} behind the SrK table
  std::vector<const VoxSegD*> seg_tab(B);
  std::vector<int*> ovf_tab(B);
  for (int b = 0; b < B; b++) {
    seg_tab[b] = host[b].segs;
    ovf_tab[b] = host[b].meta + SRM_VOX_OVERFLOW;
  }
  char* extra = (char*)tab.p + (((size_t)B * sizeof(SrK) + 15) & ~(size_t)15);
  LG_CHECK(cudaMemcpyAsync(extra, seg_tab.data(), (size_t)B * sizeof(void*), cudaMemcpyHostToDevice, st));
  LG_CHECK(cudaMemcpyAsync(extra + (size_t)B * sizeof(void*), ovf_tab.data(), (size_t)B * sizeof(void*), cudaMemcpyHostToDevice, st));
  int rc = lg_vox_small_batch((const VoxSegD* const*)extra, (int* const*)(extra + (size_t)B * sizeof(void*)), max_R, B, small_rings ? 4096 : 16384, st,
                              launches);
  if (rc) return rc;
  sr_concat_batch_kernel<<<dim3(max_R, B), 256, 0, st>>>(d_tab);
  (*launches)++;
  LG_CHECK(cudaGetLastError());
  LG_CHECK(cudaStreamSynchronize(st));  // the tables above are host vectors
  return LOAM_OK;
}

#ifdef LG_SEL_DEBUG
extern "C" int loam_debug_sel(long long* out8, int clear) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out8, g_sel_dbg, sizeof(long long) * 8);
  if (clear) {
    long long z[8] = {0};
    cudaMemcpyToSymbol(g_sel_dbg, z, sizeof(z));
  }
  return 0;
}
#endif
