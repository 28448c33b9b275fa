// K5 — voxel-grid centroid down-sampling, the B200 replacement for pcl::VoxelGrid<PointXYZI>::filter at the
// reference's call sites SR:677-683 (0.2 m per ring), LM:736-744 (0.2 / 0.4 m stacks), LM:1061-1079 (per valid
// cube) and LM:1092-1094 (surround).  Semantics (PCL 1.8.0): cell = floor(p * (1/leaf)) - min_b,
// linear id x-fastest, one output per occupied cell in ascending id, centroid of all four fields, in-cell sums taken
// in ascending input index (our deterministic choice), index-overflow => input returned unfiltered.
//
// Two paths, both batched over independent SEGMENTS (rings / stacks / map cubes):
//   small: one CTA per segment, keys (cell << 32 | index) bitonic-sorted in shared memory (<= 16384 points);
//   big  : global LSD radix sort of (segment << 32 | cell, index), 8 bits per pass, any size.
// HBM roofline: algorithmic bytes 16 (M + V) per call.
#include <cooperative_groups.h>
#include <float.h>

#include "lg_voxel.h"

namespace {

__device__ __forceinline__ int f2ord(float f) {
  int b = __float_as_int(f);
  return b >= 0 ? b : b ^ 0x7fffffff;
}
__device__ __forceinline__ float ord2f(int o) { return __int_as_float(o >= 0 ? o : o ^ 0x7fffffff); }

// Shared by both paths: PCL's grid set-up from the float bounding box.  Returns false on index overflow.
struct VoxGrid {
  int minb0, minb1, minb2, divx, divxy;
  float inv;
};
__device__ __forceinline__ bool vox_grid_setup(float mnx, float mny, float mnz, float mxx, float mxy, float mxz, float leaf, VoxGrid& g) {
  float inv = 1.0f / leaf;
  g.inv = inv;
  long long dx = (long long)((mxx - mnx) * inv) + 1;
  long long dy = (long long)((mxy - mny) * inv) + 1;
  long long dz = (long long)((mxz - mnz) * inv) + 1;
  if (dx * dy * dz > 2147483647ll) return false;
  g.minb0 = (int)floorf(mnx * inv);
  g.minb1 = (int)floorf(mny * inv);
  g.minb2 = (int)floorf(mnz * inv);
  int maxb0 = (int)floorf(mxx * inv), maxb1 = (int)floorf(mxy * inv);
  g.divx = maxb0 - g.minb0 + 1;
  g.divxy = g.divx * (maxb1 - g.minb1 + 1);
  return true;
}
__device__ __forceinline__ int vox_cell(const VoxGrid& g, float4 p) {
  int i0 = (int)(floorf(p.x * g.inv) - (float)g.minb0);
  int i1 = (int)(floorf(p.y * g.inv) - (float)g.minb1);
  int i2 = (int)(floorf(p.z * g.inv) - (float)g.minb2);
  return i0 + i1 * g.divx + i2 * g.divxy;
}

// ------------------------------------------------------------------------------------------------ small path
template <int CAP, int NT>
__device__ __forceinline__ void vox_small_body(const VoxSegD* __restrict__ segs, int* __restrict__ overflow) {
  extern __shared__ unsigned long long skeys[];
  __shared__ float s_red[6][NT / 32];
  __shared__ int s_cnt[NT / 32];
  __shared__ int s_scan[NT / 32 + 2];
  __shared__ float s_bb[6];
  __shared__ int s_nvalid;
  const VoxSegD sg = segs[blockIdx.x];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int n = sg.n;
  if (n > CAP) {
    if (tid == 0) {
      atomicExch(overflow, 1);
      *sg.out_count = -1;
    }
    return;
  }
  if (n <= 0) {
    if (tid == 0) *sg.out_count = 0;
    return;
  }
  // small segments keep their points in shared memory: the segment is read from global memory once (the bounding-box
  // pass), the key pass and the centroid gather -- a serial chain of dependent loads per voxel -- read the staged copy
  constexpr bool STAGE = CAP <= 4096;
  float4* s_pts = reinterpret_cast<float4*>(skeys + CAP);                      // [CAP] when STAGE
  unsigned char* s_take = reinterpret_cast<unsigned char*>(s_pts + CAP);       // [CAP] when STAGE
  float mn0 = FLT_MAX, mn1 = FLT_MAX, mn2 = FLT_MAX, mx0 = -FLT_MAX, mx1 = -FLT_MAX, mx2 = -FLT_MAX;
  int cnt = 0;
  for (int i0 = tid; i0 < n; i0 += 4 * NT) {  // four loads in flight per thread
    float4 p[4];
    bool take[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int i = i0 + u * NT;
      take[u] = i < n && (!sg.valid || sg.valid[i]);
      if (i < n) p[u] = sg.in[i];
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      if (STAGE && i0 + u * NT < n) {
        s_pts[i0 + u * NT] = p[u];
        s_take[i0 + u * NT] = take[u];
      }
      if (take[u]) {
        mn0 = fminf(mn0, p[u].x); mn1 = fminf(mn1, p[u].y); mn2 = fminf(mn2, p[u].z);
        mx0 = fmaxf(mx0, p[u].x); mx1 = fmaxf(mx1, p[u].y); mx2 = fmaxf(mx2, p[u].z);
        cnt++;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn0 = fminf(mn0, __shfl_xor_sync(0xffffffffu, mn0, o)); mn1 = fminf(mn1, __shfl_xor_sync(0xffffffffu, mn1, o));
    mn2 = fminf(mn2, __shfl_xor_sync(0xffffffffu, mn2, o)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, o));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, o)); mx2 = fmaxf(mx2, __shfl_xor_sync(0xffffffffu, mx2, o));
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if (lane == 0) {
    s_red[0][w] = mn0; s_red[1][w] = mn1; s_red[2][w] = mn2; s_red[3][w] = mx0; s_red[4][w] = mx1; s_red[5][w] = mx2;
    s_cnt[w] = cnt;
  }
  __syncthreads();
  if (tid == 0) {
    for (int k = 1; k < NT / 32; k++) {
      s_red[0][0] = fminf(s_red[0][0], s_red[0][k]); s_red[1][0] = fminf(s_red[1][0], s_red[1][k]);
      s_red[2][0] = fminf(s_red[2][0], s_red[2][k]); s_red[3][0] = fmaxf(s_red[3][0], s_red[3][k]);
      s_red[4][0] = fmaxf(s_red[4][0], s_red[4][k]); s_red[5][0] = fmaxf(s_red[5][0], s_red[5][k]);
      s_cnt[0] += s_cnt[k];
    }
    for (int k = 0; k < 6; k++) s_bb[k] = s_red[k][0];
    s_nvalid = s_cnt[0];
  }
  __syncthreads();
  const int nvalid = s_nvalid;
  if (nvalid == 0) {
    if (tid == 0) *sg.out_count = 0;
    return;
  }
  VoxGrid g;
  const bool ok = vox_grid_setup(s_bb[0], s_bb[1], s_bb[2], s_bb[3], s_bb[4], s_bb[5], sg.leaf, g);
  int P = 2;
  while (P < n) P <<= 1;
#pragma unroll 4
  for (int i = tid; i < P; i += NT) {
    unsigned long long key = ~0ull;
    if (STAGE) {
      if (i < n && s_take[i]) {
        int cell = ok ? vox_cell(g, s_pts[i]) : i;
        key = ((unsigned long long)(unsigned int)cell << 32) | (unsigned int)i;
      }
    } else if (i < n && (!sg.valid || sg.valid[i])) {
      int cell = ok ? vox_cell(g, sg.in[i]) : i;
      key = ((unsigned long long)(unsigned int)cell << 32) | (unsigned int)i;
    }
    skeys[i] = key;
  }
  __syncthreads();
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < (P >> 1); t += NT) {
        int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        int l = i | j;
        unsigned long long a = skeys[i], b = skeys[l];
        bool asc = (i & k) == 0;
        if ((a > b) == asc) {
          skeys[i] = b;
          skeys[l] = a;
        }
      }
      __syncthreads();
    }
  }
  const int per = (nvalid + NT - 1) / NT;
  const int b = min(tid * per, nvalid), e = min(b + per, nvalid);
  int local = 0;
  for (int i = b; i < e; i++) local += (i == 0 || (unsigned int)(skeys[i] >> 32) != (unsigned int)(skeys[i - 1] >> 32)) ? 1 : 0;
  int V;
  int r = block_excl_scan<NT>(local, &V, s_scan);
  for (int i = b; i < e; i++) {
    unsigned int cell = (unsigned int)(skeys[i] >> 32);
    if (i == 0 || cell != (unsigned int)(skeys[i - 1] >> 32)) {
      float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
      int j = i;
      while (j < nvalid && (unsigned int)(skeys[j] >> 32) == cell) {
        const unsigned int src = (unsigned int)(skeys[j] & 0xffffffffull);
        float4 p = STAGE ? s_pts[src] : sg.in[src];
        sx = sx + p.x; sy = sy + p.y; sz = sz + p.z; si = si + p.w;
        j++;
      }
      float c = (float)(j - i);
      sg.out[r++] = make_float4(sx / c, sy / c, sz / c, si / c);
    }
  }
  if (tid == 0) *sg.out_count = V;
}
// dynamic shared memory of vox_small_kernel<CAP, NT>: keys, and for small segments the staged points + validity flags
template <int CAP>
constexpr int vox_small_smem() { return CAP * 8 + (CAP <= 4096 ? CAP * 17 : 0); }


// ------------------------------------------------------------------------------------------------ split path
template <int CAP, int NT>
__global__ void __launch_bounds__(NT) vox_small_kernel(const VoxSegD* __restrict__ segs, int* __restrict__ overflow) {
  vox_small_body<CAP, NT>(segs, overflow);
}
// several segment arrays in one launch: blockIdx.y picks the array (one per sequence of a batched extraction)
template <int CAP, int NT>
__global__ void __launch_bounds__(NT) vox_small_batch_kernel(const VoxSegD* const* __restrict__ seg_tab, int* const* __restrict__ overflow_tab) {
  vox_small_body<CAP, NT>(seg_tab[blockIdx.y], overflow_tab[blockIdx.y]);
}

// Medium segments (a sweep-sized stack, 4 k .. 64 k points): the cell-id range of the segment is cut into `gridDim.x`
// equal sub-ranges, one CTA each.  Every CTA scans the whole segment (L2-resident), keeps the points of its
// sub-range, sorts them in shared memory and writes their centroids to its staging slice; a second kernel
// concatenates the slices in range order, which is ascending cell id.  Same results as vox_small_kernel.
constexpr int VS_CAP = 4096, VS_NT = 1024, VS_MLP = 4, VS_BINS = 1024;  // VS_MLP loads in flight per thread: every CTA streams the whole segment
__global__ void __launch_bounds__(VS_NT) vox_split_kernel(const VoxSegD* __restrict__ segs, float4* __restrict__ staging,
                                                           int* __restrict__ range_counts, int* __restrict__ overflow) {
  __shared__ unsigned long long skeys[VS_CAP];
  __shared__ float s_red[6][VS_NT / 32];
  __shared__ int s_scan[VS_NT / 32 + 2];
  __shared__ float s_bb[6];
  __shared__ int s_n;
  __shared__ int s_hist[VS_BINS];
  const VoxSegD sg = segs[blockIdx.y];
#ifdef VS_DEBUG
  const long long dbg_t0 = clock64();
#endif
  const int C = gridDim.x, c = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  int* my_count = range_counts + blockIdx.y * C + c;
  float4* my_out = staging + ((size_t)blockIdx.y * C + c) * VS_CAP;
  const int n = sg.n;
  if (n <= 0) {
    if (tid == 0) *my_count = 0;
    return;
  }
  float mn0 = FLT_MAX, mn1 = FLT_MAX, mn2 = FLT_MAX, mx0 = -FLT_MAX, mx1 = -FLT_MAX, mx2 = -FLT_MAX;
  for (int i0 = tid; i0 < n; i0 += VS_MLP * VS_NT) {
    float4 p[VS_MLP];
#pragma unroll
    for (int u = 0; u < VS_MLP; u++)
      if (i0 + u * VS_NT < n) p[u] = sg.in[i0 + u * VS_NT];
#pragma unroll
    for (int u = 0; u < VS_MLP; u++)
      if (i0 + u * VS_NT < n) {
        mn0 = fminf(mn0, p[u].x); mn1 = fminf(mn1, p[u].y); mn2 = fminf(mn2, p[u].z);
        mx0 = fmaxf(mx0, p[u].x); mx1 = fmaxf(mx1, p[u].y); mx2 = fmaxf(mx2, p[u].z);
      }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn0 = fminf(mn0, __shfl_xor_sync(0xffffffffu, mn0, o)); mn1 = fminf(mn1, __shfl_xor_sync(0xffffffffu, mn1, o));
    mn2 = fminf(mn2, __shfl_xor_sync(0xffffffffu, mn2, o)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, o));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, o)); mx2 = fmaxf(mx2, __shfl_xor_sync(0xffffffffu, mx2, o));
  }
  if (lane == 0) {
    s_red[0][w] = mn0; s_red[1][w] = mn1; s_red[2][w] = mn2; s_red[3][w] = mx0; s_red[4][w] = mx1; s_red[5][w] = mx2;
  }
  if (tid == 0) s_n = 0;
  __syncthreads();
  if (tid == 0) {
    for (int k = 1; k < VS_NT / 32; k++) {
      s_red[0][0] = fminf(s_red[0][0], s_red[0][k]); s_red[1][0] = fminf(s_red[1][0], s_red[1][k]);
      s_red[2][0] = fminf(s_red[2][0], s_red[2][k]); s_red[3][0] = fmaxf(s_red[3][0], s_red[3][k]);
      s_red[4][0] = fmaxf(s_red[4][0], s_red[4][k]); s_red[5][0] = fmaxf(s_red[5][0], s_red[5][k]);
    }
    for (int k = 0; k < 6; k++) s_bb[k] = s_red[k][0];
  }
  __syncthreads();
  VoxGrid g;
  const bool ok = vox_grid_setup(s_bb[0], s_bb[1], s_bb[2], s_bb[3], s_bb[4], s_bb[5], sg.leaf, g);
  if (!ok) {  // PCL would hand the input back unfiltered: leave that to the single-CTA / big paths
    if (tid == 0) {
      atomicExch(overflow, 1);
      *my_count = 0;
    }
    return;
  }
  const int maxb2 = (int)floorf(s_bb[5] * g.inv);
  const long long total = (long long)g.divxy * (maxb2 - g.minb2 + 1);
  // Split the cell-id range so that every CTA gets about n / C POINTS (an even split of the id range gave a few CTAs
  // most of the cloud and 4096-key sorts while the rest idled): coarse histogram over VS_BINS id bins, built by every
  // CTA for itself (integer counts: identical everywhere), CTA c takes the bins between the c-th and (c+1)-th
  // quantile.  Bin edges are cell-id multiples, so no voxel is ever split and the concatenation stays in id order.
  int bin_shift = 0;  // power-of-two bin width: cell >> bin_shift < VS_BINS
  while (((total - 1) >> bin_shift) >= VS_BINS) bin_shift++;
  const long long binw = 1ll << bin_shift;
  for (int b = tid; b < VS_BINS; b += VS_NT) s_hist[b] = 0;
  __syncthreads();
  for (int i0 = tid; i0 < n; i0 += VS_MLP * VS_NT) {
    float4 p[VS_MLP];
#pragma unroll
    for (int u = 0; u < VS_MLP; u++)
      if (i0 + u * VS_NT < n) p[u] = sg.in[i0 + u * VS_NT];
#pragma unroll
    for (int u = 0; u < VS_MLP; u++)
      if (i0 + u * VS_NT < n) atomicAdd(&s_hist[vox_cell(g, p[u]) >> bin_shift], 1);
  }
  __syncthreads();
  {  // inclusive prefix over the bins (VS_BINS / VS_NT consecutive bins per thread)
    constexpr int PER = VS_BINS / VS_NT;
    int v[PER], local = 0;
#pragma unroll
    for (int k = 0; k < PER; k++) {
      v[k] = s_hist[tid * PER + k];
      local += v[k];
    }
    int tot;
    int run = block_excl_scan<VS_NT>(local, &tot, s_scan);
#pragma unroll
    for (int k = 0; k < PER; k++) {
      run += v[k];
      s_hist[tid * PER + k] = run;
    }
  }
  __syncthreads();
  // first bin whose inclusive prefix exceeds the quantile target = start of CTA c's range (monotone in c)
  auto first_bin = [&](int cc) {
    if (cc <= 0) return 0;
    if (cc >= C) return VS_BINS;
    const long long target = (long long)n * cc / C;
    int lo_b = 0, hi_b = VS_BINS;
    while (lo_b < hi_b) {
      const int mid = (lo_b + hi_b) >> 1;
      if ((long long)s_hist[mid] > target) hi_b = mid; else lo_b = mid + 1;
    }
    return lo_b;
  };
  const long long lo = (long long)first_bin(c) * binw, hi = (c + 1 >= C) ? total + binw : (long long)first_bin(c + 1) * binw;
  for (int i0 = tid; i0 < n; i0 += VS_MLP * VS_NT) {
    float4 p[VS_MLP];
#pragma unroll
    for (int u = 0; u < VS_MLP; u++)
      if (i0 + u * VS_NT < n) p[u] = sg.in[i0 + u * VS_NT];
#pragma unroll
    for (int u = 0; u < VS_MLP; u++) {
      const int i = i0 + u * VS_NT;
      if (i < n) {
        int cell = vox_cell(g, p[u]);
        if (cell >= lo && cell < hi) {
          int pos = atomicAdd(&s_n, 1);  // slot order is irrelevant: the keys embed the point index and are sorted next
          if (pos < VS_CAP) skeys[pos] = ((unsigned long long)(unsigned int)cell << 32) | (unsigned int)i;
        }
      }
    }
  }
  __syncthreads();
  const int m = s_n;
  if (m > VS_CAP) {
    if (tid == 0) {
      atomicExch(overflow, 1);
      *my_count = 0;
    }
    return;
  }
  if (m == 0) {
    if (tid == 0) *my_count = 0;
    return;
  }
  int P = 2;
  while (P < m) P <<= 1;
  for (int i = m + tid; i < P; i += VS_NT) skeys[i] = ~0ull;
  __syncthreads();
  for (int k = 2; k <= P; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < (P >> 1); t += VS_NT) {
        int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        int l = i | j;
        unsigned long long a = skeys[i], b = skeys[l];
        bool asc = (i & k) == 0;
        if ((a > b) == asc) {
          skeys[i] = b;
          skeys[l] = a;
        }
      }
      __syncthreads();
    }
  const int per = (m + VS_NT - 1) / VS_NT;
  const int b0 = min(tid * per, m), e0 = min(b0 + per, m);
  int local = 0;
  for (int i = b0; i < e0; i++) local += (i == 0 || (unsigned int)(skeys[i] >> 32) != (unsigned int)(skeys[i - 1] >> 32)) ? 1 : 0;
  int V;
  int r = block_excl_scan<VS_NT>(local, &V, s_scan);
  for (int i = b0; i < e0; i++) {
    unsigned int cell = (unsigned int)(skeys[i] >> 32);
    if (i == 0 || cell != (unsigned int)(skeys[i - 1] >> 32)) {
      float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
      int j = i;
      while (j < m && (unsigned int)(skeys[j] >> 32) == cell) {
        float4 p = sg.in[(unsigned int)(skeys[j] & 0xffffffffull)];
        sx = sx + p.x; sy = sy + p.y; sz = sz + p.z; si = si + p.w;
        j++;
      }
      float cc = (float)(j - i);
      my_out[r++] = make_float4(sx / cc, sy / cc, sz / cc, si / cc);
    }
  }
  if (tid == 0) *my_count = V;
#ifdef VS_DEBUG
  if (tid == 0) printf("vs seg %d cta %d n %d m %d P %d V %d cycles %lld\n", (int)blockIdx.y, c, n, m, P, V, clock64() - dbg_t0);
#endif
}

__global__ void __launch_bounds__(256) vox_split_concat_kernel(const VoxSegD* __restrict__ segs, const float4* __restrict__ staging,
                                                                const int* __restrict__ range_counts, int C) {
  const int s = blockIdx.y, c = blockIdx.x;
  const int* cnt = range_counts + s * C;
  int off = 0, tot = 0;
  for (int k = 0; k < C; k++) {
    if (k < c) off += cnt[k];
    tot += cnt[k];
  }
  const float4* src = staging + ((size_t)s * C + c) * VS_CAP;
  float4* dst = segs[s].out + off;
  for (int i = threadIdx.x; i < cnt[c]; i += blockDim.x) dst[i] = src[i];
  if (c == 0 && threadIdx.x == 0) *segs[s].out_count = tot;
}

// ------------------------------------------------------------------------------------------------ radix sort
// Lanes of the warp holding the same 8-bit digit, from eight ballots (MATCH.ANY resolves one distinct value per
// iteration and made a digit pass over 10 k keys cost ~13 us).  Invalid lanes get an unspecified mask.
__device__ __forceinline__ unsigned int lg_match8(unsigned int d, bool valid) {
  unsigned int m = __ballot_sync(0xffffffffu, valid);
#pragma unroll
  for (int b = 0; b < 8; b++) {
    const bool bit = (d >> b) & 1u;
    const unsigned int bal = __ballot_sync(0xffffffffu, bit);
    m &= bit ? bal : ~bal;
  }
  return m;
}

constexpr int RS_NT = 256, RS_ITEMS = 8, RS_TILE = RS_NT * RS_ITEMS;

__global__ void __launch_bounds__(RS_NT) rs_hist_kernel(const unsigned long long* __restrict__ keys, int n, int shift,
                                                         unsigned int* __restrict__ hist, int nblocks) {
  __shared__ unsigned int h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  int base = blockIdx.x * RS_TILE;
#pragma unroll
  for (int c = 0; c < RS_ITEMS; c++) {
    int i = base + c * RS_NT + threadIdx.x;
    if (i < n) atomicAdd(&h[(unsigned int)(keys[i] >> shift) & 255u], 1u);
  }
  __syncthreads();
  hist[threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

// in-place exclusive scan of `total` counters by one CTA
__global__ void __launch_bounds__(1024) rs_scan_kernel(unsigned int* __restrict__ a, int total) {
  __shared__ int s_scan[34];
  int per = (total + 1023) / 1024;
  int b = min((int)threadIdx.x * per, total), e = min(b + per, total);
  int local = 0;
  for (int i = b; i < e; i++) local += (int)a[i];
  int tot;
  int r = block_excl_scan<1024>(local, &tot, s_scan);
  for (int i = b; i < e; i++) {
    int v = (int)a[i];
    a[i] = (unsigned int)r;
    r += v;
  }
}

__global__ void __launch_bounds__(RS_NT) rs_scatter_kernel(const unsigned long long* __restrict__ kin, const unsigned int* __restrict__ vin,
                                                            unsigned long long* __restrict__ kout, unsigned int* __restrict__ vout, int n,
                                                            int shift, const unsigned int* __restrict__ hist, int nblocks) {
  __shared__ unsigned int s_base[256];
  __shared__ unsigned int s_wcnt[RS_NT / 32][256];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  s_base[tid] = hist[tid * nblocks + blockIdx.x];
  int base = blockIdx.x * RS_TILE;
  for (int c = 0; c < RS_ITEMS; c++) {
#pragma unroll
    for (int k = 0; k < RS_NT / 32; k++) s_wcnt[k][tid] = 0;
    __syncthreads();
    int i = base + c * RS_NT + tid;
    bool valid = i < n;
    unsigned long long key = valid ? kin[i] : 0ull;
    unsigned int d = (unsigned int)(key >> shift) & 255u;
    unsigned int m = lg_match8(d, valid);
    unsigned int rank = __popc(m & ((1u << lane) - 1u));
    if (valid && rank == 0) s_wcnt[w][d] = __popc(m);
    __syncthreads();
    if (valid) {
      unsigned int off = s_base[d] + rank;
      for (int k = 0; k < w; k++) off += s_wcnt[k][d];
      kout[off] = key;
      vout[off] = vin[i];
    }
    __syncthreads();
    unsigned int add = 0;
#pragma unroll
    for (int k = 0; k < RS_NT / 32; k++) add += s_wcnt[k][tid];
    s_base[tid] += add;
    __syncthreads();
  }
}

// Whole stable LSD radix sort of up to 16384 (key, value) pairs in ONE launch of ONE 8-CTA cluster: the map stage sorts
// ~10 k new points twice per run (by cube, then by voxel), and at that size the three-kernels-per-digit path above is
// pure launch latency (~23 us per digit) while a single CTA is bound by one SM's issue rate (~10 us per digit).
// Every CTA keeps ALL keys in shared memory (they never move) and owns one eighth of the positions of the current
// order; what moves is a 16-bit permutation in global memory.  A digit pass: rank every element inside its warp's
// rows with ballots (stable: row, then lane) -> per-CTA (digit, warp) prefix + digit totals -> cluster barrier ->
// every CTA reads its peers' digit totals out of their shared memory (DSMEM) and derives its global bases ->
// scatter the permutation -> cluster barrier.
constexpr int RSS_CAP = 16384, RSS_NT = 1024, RSC_CTAS = 8;
constexpr int RSS_ROWS = RSS_CAP / RSC_CTAS / RSS_NT;  // rows per warp (2)
constexpr int RSS_CW = 258;  // counter row pitch (warp-major, padded: the 32 lanes of a warp hit 32 different digits)
constexpr int RSS_SMEM = RSS_CAP * 8 + 32 * RSS_CW * 2;
struct RsShifts {
  int s[8];
  int n;
};
__global__ void __cluster_dims__(RSC_CTAS, 1, 1) __launch_bounds__(RSS_NT)
    rs_cluster_kernel(const unsigned long long* __restrict__ kin, const unsigned int* __restrict__ vin, unsigned long long* __restrict__ kout,
                      unsigned int* __restrict__ vout, int n, RsShifts sh, unsigned short* perm /* [2][RSS_CAP] */) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ unsigned long long rss_smem[];
  unsigned long long* s_key = rss_smem;
  unsigned short* s_cnt = reinterpret_cast<unsigned short*>(s_key + RSS_CAP);  // [32 warps][RSS_CW]
  __shared__ unsigned int s_tot[256], s_base[256];
  __shared__ int s_scan[34];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int c = (int)cluster.block_rank();
  const int chunk = ((n + RSC_CTAS * 32 - 1) / (RSC_CTAS * 32)) * 32;  // positions per CTA
  const int p0 = c * chunk, p1 = min(n, p0 + chunk);
  const int R = (chunk / 32 + 31) / 32;  // rows per warp
  unsigned short* in = perm;
  unsigned short* out = perm + RSS_CAP;
  for (int i = tid; i < n; i += RSS_NT) s_key[i] = kin[i];
  for (int p = p0 + tid; p < p1; p += RSS_NT) in[p] = (unsigned short)p;
  __syncthreads();
  for (int pass = 0; pass < sh.n; pass++) {
    const int shift = sh.s[pass];
    for (int i = tid; i < 32 * RSS_CW; i += RSS_NT) s_cnt[i] = 0;
    __syncthreads();
    unsigned int e[RSS_ROWS], rk[RSS_ROWS];
#pragma unroll
    for (int r = 0; r < RSS_ROWS; r++) {
      e[r] = rk[r] = 0;
      if (r < R) {
        const int p = p0 + (w * R + r) * 32 + lane;
        const bool valid = p < p1;
        const unsigned int ei = valid ? __ldcg(&in[p]) : 0u;  // written by other SMs in the previous pass: bypass L1
        const unsigned int d = valid ? ((unsigned int)(s_key[ei] >> shift) & 255u) : 0u;
        const unsigned int m = lg_match8(d, valid);
        const unsigned int rank = __popc(m & ((1u << lane) - 1u));
        unsigned int before = 0;
        if (valid) before = s_cnt[w * RSS_CW + d];
        __syncwarp();
        if (valid && rank == 0) s_cnt[w * RSS_CW + d] = (unsigned short)(before + __popc(m));
        __syncwarp();
        e[r] = ei;
        rk[r] = before + rank;
      }
    }
    __syncthreads();
    {  // (digit, warp) exclusive prefix inside the CTA: thread t owns digit t / 4, warps (t % 4) * 8 .. + 7
      const int dd = tid >> 2, w0 = (tid & 3) * 8;
      int v[8], local = 0;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        v[k] = s_cnt[(w0 + k) * RSS_CW + dd];
        local += v[k];
      }
      int incl = local;
      int t1 = __shfl_up_sync(0xffffffffu, incl, 1, 4);
      if ((lane & 3) >= 1) incl += t1;
      int t2 = __shfl_up_sync(0xffffffffu, incl, 2, 4);
      if ((lane & 3) >= 2) incl += t2;
      const int total = __shfl_sync(0xffffffffu, incl, 3, 4);
      int run = incl - local;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        s_cnt[(w0 + k) * RSS_CW + dd] = (unsigned short)run;
        run += v[k];
      }
      if ((tid & 3) == 0) s_tot[dd] = (unsigned int)total;
    }
    cluster.sync();
    {  // global base of (digit, this CTA): all smaller digits everywhere + the same digit in lower-ranked CTAs
      unsigned int before = 0, all = 0;
      if (tid < 256) {
#pragma unroll
        for (int cc = 0; cc < RSC_CTAS; cc++) {
          const unsigned int t = cluster.map_shared_rank(s_tot, cc)[tid];
          if (cc < c) before += t;
          all += t;
        }
      }
      int tot;
      const int ex = block_excl_scan<RSS_NT>((int)all, &tot, s_scan);
      if (tid < 256) s_base[tid] = (unsigned int)ex + before;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RSS_ROWS; r++) {
      if (r < R) {
        const int p = p0 + (w * R + r) * 32 + lane;
        if (p < p1) {
          const unsigned int d = (unsigned int)(s_key[e[r]] >> shift) & 255u;
          out[s_base[d] + s_cnt[w * RSS_CW + d] + rk[r]] = (unsigned short)e[r];
        }
      }
    }
    cluster.sync();  // the new order is complete (and nobody still reads this pass's totals) before the next pass
    unsigned short* tmp = in;
    in = out;
    out = tmp;
  }
  for (int p = p0 + tid; p < p1; p += RSS_NT) {
    const unsigned int ei = __ldcg(&in[p]);
    kout[p] = s_key[ei];
    vout[p] = vin[ei];
  }
}

// ------------------------------------------------------------------------------------------------ big path
__device__ __forceinline__ int find_seg(const int* __restrict__ seg_off, int nseg, int i) {
  int lo = 0, hi = nseg;  // largest s with seg_off[s] <= i
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (seg_off[mid] <= i) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void vb_init_kernel(int* __restrict__ bb, int nseg, int* __restrict__ out_start, int* __restrict__ out_end) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nseg * 6) bb[i] = (i % 6) < 3 ? 0x7fffffff : (int)0x80000000;
  if (i < nseg) {
    out_start[i] = 0;
    out_end[i] = 0;
  }
}

__global__ void vb_bbox_kernel(const float4* __restrict__ in, const int* __restrict__ seg_off, int nseg, int M, int* __restrict__ bb) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int seg = -1;
  int o[6];
  if (i < M) {
    seg = find_seg(seg_off, nseg, i);
    float4 p = in[i];
    o[0] = f2ord(p.x); o[1] = f2ord(p.y); o[2] = f2ord(p.z);
    o[3] = o[0]; o[4] = o[1]; o[5] = o[2];
  }
  unsigned int m = __match_any_sync(0xffffffffu, seg);
  if (seg < 0) return;
  // reduce inside the group of lanes that share a segment (usually the whole warp)
  int lane = threadIdx.x & 31;
  int leader = __ffs(m) - 1;
  if (m == 0xffffffffu) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
      for (int k = 0; k < 3; k++) {
        o[k] = min(o[k], __shfl_xor_sync(0xffffffffu, o[k], off));
        o[k + 3] = max(o[k + 3], __shfl_xor_sync(0xffffffffu, o[k + 3], off));
      }
    }
    if (lane == leader) {
#pragma unroll
      for (int k = 0; k < 3; k++) {
        atomicMin(&bb[seg * 6 + k], o[k]);
        atomicMax(&bb[seg * 6 + 3 + k], o[k + 3]);
      }
    }
  } else {
#pragma unroll
    for (int k = 0; k < 3; k++) {
      atomicMin(&bb[seg * 6 + k], o[k]);
      atomicMax(&bb[seg * 6 + 3 + k], o[k + 3]);
    }
  }
}

__global__ void vb_key_kernel(const float4* __restrict__ in, const int* __restrict__ seg_off, const float* __restrict__ seg_leaf, int nseg,
                              int M, const int* __restrict__ bb, unsigned long long* __restrict__ keys, unsigned int* __restrict__ vals) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  int seg = find_seg(seg_off, nseg, i);
  const int* b = bb + seg * 6;
  VoxGrid g;
  bool ok = vox_grid_setup(ord2f(b[0]), ord2f(b[1]), ord2f(b[2]), ord2f(b[3]), ord2f(b[4]), ord2f(b[5]), seg_leaf[seg], g);
  int cell = ok ? vox_cell(g, in[i]) : (i - seg_off[seg]);
  keys[i] = ((unsigned long long)(unsigned int)seg << 32) | (unsigned int)cell;
  vals[i] = (unsigned int)i;
}


// ------------------------------------------------------------------------------------------------ merge path
// Voxel-gridding a cube whose OLD cloud is the output of the previous voxel grid (one point per cell, ascending
// cell id) plus a few NEW points: only the new points are sorted; the two sorted key sequences are merged by rank
// (old before new on equal keys = ascending input index, exactly the order the full sort would produce) and the
// usual head / centroid kernels run on the merged sequence.  The old part is CHECKED to be strictly ascending under
// the new bounding box (a centroid can drift across a cell face by an ulp); if not, the caller falls back to the
// full sort.  Keys: (segment << 32) | cell with cell < 2^24 (cube-sized segments), so the radix sort of the new part
// skips bits 24..31.
__global__ void vm_key_kernel(const float4* __restrict__ in, const int* __restrict__ seg_off, const float* __restrict__ seg_leaf, int nseg,
                              int n, int val_base, const int* __restrict__ bb, unsigned long long* __restrict__ keys,
                              unsigned int* __restrict__ vals, int check_sorted, int* __restrict__ flags) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int seg = find_seg(seg_off, nseg, i);
  const int* b = bb + seg * 6;
  VoxGrid g;
  bool ok = vox_grid_setup(ord2f(b[0]), ord2f(b[1]), ord2f(b[2]), ord2f(b[3]), ord2f(b[4]), ord2f(b[5]), seg_leaf[seg], g);
  int cell = ok ? vox_cell(g, in[i]) : 0;
  if (!ok || cell >= (1 << 24) || cell < 0) atomicOr(&flags[0], 1);  // not a cube-sized segment: full path needed
  unsigned long long key = ((unsigned long long)(unsigned int)seg << 32) | (unsigned int)cell;
  keys[i] = key;
  vals[i] = (unsigned int)(val_base + i);
  if (check_sorted && i > seg_off[seg]) {  // strictly ascending inside the segment (segments ascend by construction)
    int pc = vox_cell(g, in[i - 1]);
    if (pc >= cell) atomicOr(&flags[0], 2);
  }
}

__global__ void vm_merge_kernel(const unsigned long long* __restrict__ ko, const unsigned int* __restrict__ vo, int n_old,
                                const unsigned long long* __restrict__ kn, const unsigned int* __restrict__ vn, int n_new,
                                unsigned long long* __restrict__ km, unsigned int* __restrict__ vm) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_old) {
    unsigned long long key = ko[i];
    int lo = 0, hi = n_new;  // number of new keys < key
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (kn[mid] < key) lo = mid + 1; else hi = mid;
    }
    km[i + lo] = key;
    vm[i + lo] = vo[i];
  } else if (i < n_old + n_new) {
    int j = i - n_old;
    unsigned long long key = kn[j];
    int lo = 0, hi = n_old;  // number of old keys <= key
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (ko[mid] <= key) lo = mid + 1; else hi = mid;
    }
    km[j + lo] = key;
    vm[j + lo] = vn[j];
  }
}

constexpr int VC_NT = 256, VC_ITEMS = 4, VC_TILE = VC_NT * VC_ITEMS;

__global__ void __launch_bounds__(VC_NT) vb_count_kernel(const unsigned long long* __restrict__ keys, int M, unsigned int* __restrict__ block_sums) {
  __shared__ int s_w[VC_NT / 32];
  int base = blockIdx.x * VC_TILE + threadIdx.x * VC_ITEMS;
  int c = 0;
#pragma unroll
  for (int k = 0; k < VC_ITEMS; k++) {
    int i = base + k;
    if (i < M) c += (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int k = 0; k < VC_NT / 32; k++) t += s_w[k];
    block_sums[blockIdx.x] = (unsigned int)t;
  }
}

__global__ void __launch_bounds__(VC_NT) vb_centroid_kernel(const float4* __restrict__ in, const unsigned long long* __restrict__ keys,
                                                             const unsigned int* __restrict__ vals, int M, const unsigned int* __restrict__ block_off,
                                                             const int* __restrict__ seg_off, float4* __restrict__ out,
                                                             int* __restrict__ out_start, int* __restrict__ out_end) {
  __shared__ int s_scan[VC_NT / 32 + 2];
  int base = blockIdx.x * VC_TILE + threadIdx.x * VC_ITEMS;
  int local = 0;
  bool head[VC_ITEMS];
#pragma unroll
  for (int k = 0; k < VC_ITEMS; k++) {
    int i = base + k;
    head[k] = (i < M) && (i == 0 || keys[i] != keys[i - 1]);
    local += head[k] ? 1 : 0;
  }
  int tot;
  int r = block_excl_scan<VC_NT>(local, &tot, s_scan) + (int)block_off[blockIdx.x];
#pragma unroll
  for (int k = 0; k < VC_ITEMS; k++) {
    int i = base + k;
    if (i >= M) break;
    unsigned long long key = keys[i];
    int seg = (int)(key >> 32);
    if (head[k]) {
      float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
      int j = i;
      while (j < M && keys[j] == key) {
        float4 p = in[vals[j]];
        sx = sx + p.x; sy = sy + p.y; sz = sz + p.z; si = si + p.w;
        j++;
      }
      float c = (float)(j - i);
      out[r] = make_float4(sx / c, sy / c, sz / c, si / c);
      if (i == seg_off[seg]) out_start[seg] = r;
      r++;
    }
    if (i == seg_off[seg + 1] - 1) out_end[seg] = r;  // r = rank of this element's cell + 1
  }
}

__global__ void gather_kernel(const CopyEnt* __restrict__ ents, int nent, float4* __restrict__ dst) {
  // blockIdx.y = entry, blockIdx.x strides over the entry's points
  for (int e = blockIdx.y; e < nent; e += gridDim.y) {
    CopyEnt ce = ents[e];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ce.n; i += gridDim.x * blockDim.x) dst[ce.dst_off + i] = ce.src[i];
  }
}

}  // namespace

// ---------------------------------------------------------------------------------------------------- host side
int lg_vox_small(const VoxSegD* d_segs, int nseg, int max_seg_hint, int* d_overflow, cudaStream_t st, long long* launches) {
  if (nseg <= 0) return LOAM_OK;
  if (max_seg_hint <= 4096) {
    static bool attr_small[64] = {};  // the opt-in above 48 KB is per device
    int dev = 0;
    LG_CHECK(cudaGetDevice(&dev));
    if (!attr_small[dev & 63]) {
      LG_CHECK(cudaFuncSetAttribute(vox_small_kernel<4096, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, vox_small_smem<4096>()));
      attr_small[dev & 63] = true;
    }
    vox_small_kernel<4096, 512><<<nseg, 512, vox_small_smem<4096>(), st>>>(d_segs, d_overflow);
  } else {
    static bool attr_set = false;
    if (!attr_set) {
      cudaFuncSetAttribute(vox_small_kernel<16384, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * (int)sizeof(unsigned long long));
      attr_set = true;
    }
    vox_small_kernel<16384, 1024><<<nseg, 1024, 16384 * sizeof(unsigned long long), st>>>(d_segs, d_overflow);
  }
  (*launches)++;
  LG_CHECK(cudaGetLastError());
  return LOAM_OK;
}


int lg_vox_small_batch(const VoxSegD* const* d_seg_tab, int* const* d_overflow_tab, int max_nseg, int B, int max_seg_hint, cudaStream_t st,
                       long long* launches) {
  if (max_nseg <= 0 || B <= 0) return LOAM_OK;
  if (max_seg_hint <= 4096) {
    static bool attr_small[64] = {};
    int dev = 0;
    LG_CHECK(cudaGetDevice(&dev));
    if (!attr_small[dev & 63]) {
      LG_CHECK(cudaFuncSetAttribute(vox_small_batch_kernel<4096, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, vox_small_smem<4096>()));
      attr_small[dev & 63] = true;
    }
    vox_small_batch_kernel<4096, 512><<<dim3(max_nseg, B), 512, vox_small_smem<4096>(), st>>>(d_seg_tab, d_overflow_tab);
  } else {
    static bool attr_set = false;
    if (!attr_set) {
      cudaFuncSetAttribute(vox_small_batch_kernel<16384, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           16384 * (int)sizeof(unsigned long long));
      attr_set = true;
    }
    vox_small_batch_kernel<16384, 1024><<<dim3(max_nseg, B), 1024, 16384 * sizeof(unsigned long long), st>>>(d_seg_tab, d_overflow_tab);
  }
  (*launches)++;
  LG_CHECK(cudaGetLastError());
  return LOAM_OK;
}

int lg_vox_split(DevBuf& staging, DevBuf& counts, const VoxSegD* d_segs, int nseg, int* d_overflow, cudaStream_t st, long long* launches) {
  if (nseg <= 0) return LOAM_OK;
  const int C = 32;
  LG_CHECK(staging.ensure((size_t)nseg * C * VS_CAP * sizeof(float4), st));
  LG_CHECK(counts.ensure((size_t)nseg * C * sizeof(int), st));
  dim3 grid(C, nseg);
  vox_split_kernel<<<grid, VS_NT, 0, st>>>(d_segs, staging.as<float4>(), counts.as<int>(), d_overflow);
  vox_split_concat_kernel<<<grid, 256, 0, st>>>(d_segs, staging.as<float4>(), counts.as<int>(), C);
  (*launches) += 2;
  LG_CHECK(cudaGetLastError());
  return LOAM_OK;
}

static int rs_small_launch(RadixWs& ws, int n, const int* shifts, int nshifts, cudaStream_t st, long long* launches, int* result_in_b) {
  static bool attr[64] = {};
  int dev = 0;
  LG_CHECK(cudaGetDevice(&dev));
  if (!attr[dev & 63]) {
    LG_CHECK(cudaFuncSetAttribute(rs_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RSS_SMEM));
    attr[dev & 63] = true;
  }
  RsShifts sh;
  sh.n = nshifts;
  for (int i = 0; i < 8; i++) sh.s[i] = i < nshifts ? shifts[i] : 0;
  LG_CHECK(ws.perm.ensure((size_t)2 * RSS_CAP * sizeof(unsigned short), st));
  rs_cluster_kernel<<<RSC_CTAS, RSS_NT, RSS_SMEM, st>>>(ws.keysA.as<unsigned long long>(), ws.valsA.as<unsigned int>(),
                                                        ws.keysB.as<unsigned long long>(), ws.valsB.as<unsigned int>(), n, sh,
                                                        ws.perm.as<unsigned short>());
  (*launches)++;
  LG_CHECK(cudaGetLastError());
  *result_in_b = 1;
  return LOAM_OK;
}

int lg_radix_sort(RadixWs& ws, int n, int bits, cudaStream_t st, long long* launches, int* result_in_b) {
  // keys/vals in ws.keysA/valsA; sorted result ends in A or B (result_in_b)
  if (n <= RSS_CAP && bits <= 64) {
    int shifts[8], ns = 0;
    for (int shift = 0; shift < bits; shift += 8) shifts[ns++] = shift;
    return rs_small_launch(ws, n, shifts, ns, st, launches, result_in_b);
  }
  int nblocks = lg_div_up(n, RS_TILE);
  LG_CHECK(ws.hist.ensure((size_t)256 * nblocks * sizeof(unsigned int), st));
  unsigned long long* ka = ws.keysA.as<unsigned long long>();
  unsigned long long* kb = ws.keysB.as<unsigned long long>();
  unsigned int* va = ws.valsA.as<unsigned int>();
  unsigned int* vb = ws.valsB.as<unsigned int>();
  int flip = 0;
  for (int shift = 0; shift < bits; shift += 8) {
    rs_hist_kernel<<<nblocks, RS_NT, 0, st>>>(ka, n, shift, ws.hist.as<unsigned int>(), nblocks);
    rs_scan_kernel<<<1, 1024, 0, st>>>(ws.hist.as<unsigned int>(), 256 * nblocks);
    rs_scatter_kernel<<<nblocks, RS_NT, 0, st>>>(ka, va, kb, vb, n, shift, ws.hist.as<unsigned int>(), nblocks);
    (*launches) += 3;
    std::swap(ka, kb);
    std::swap(va, vb);
    flip ^= 1;
  }
  LG_CHECK(cudaGetLastError());
  *result_in_b = flip;
  return LOAM_OK;
}

int lg_radix_ensure(RadixWs& ws, int n, cudaStream_t st) {
  LG_CHECK(ws.keysA.ensure((size_t)n * 8, st));
  LG_CHECK(ws.keysB.ensure((size_t)n * 8, st));
  LG_CHECK(ws.valsA.ensure((size_t)n * 4, st));
  LG_CHECK(ws.valsB.ensure((size_t)n * 4, st));
  return LOAM_OK;
}

int lg_vox_big(VoxBigWs& ws, const float4* d_in, const int* d_seg_off, const float* d_seg_leaf, int nseg, int M, float4* d_out,
               int* d_out_start, int* d_out_end, cudaStream_t st, long long* launches) {
  if (nseg <= 0) return LOAM_OK;
  LgProfScope prof_scope(LGK_VOXEL, st, (double)M);
  LG_CHECK(ws.bb.ensure((size_t)nseg * 6 * sizeof(int), st));
  vb_init_kernel<<<lg_div_up(nseg * 6, 256), 256, 0, st>>>(ws.bb.as<int>(), nseg, d_out_start, d_out_end);
  (*launches)++;
  if (M <= 0) return LOAM_OK;
  int rc = lg_radix_ensure(ws.rs, M, st);
  if (rc) return rc;
  int nb = lg_div_up(M, 256);
  vb_bbox_kernel<<<nb, 256, 0, st>>>(d_in, d_seg_off, nseg, M, ws.bb.as<int>());
  vb_key_kernel<<<nb, 256, 0, st>>>(d_in, d_seg_off, d_seg_leaf, nseg, M, ws.bb.as<int>(), ws.rs.keysA.as<unsigned long long>(),
                                    ws.rs.valsA.as<unsigned int>());
  (*launches) += 2;
  int segbits = 0;
  while ((1 << segbits) < nseg) segbits++;
  int bits = nseg > 1 ? 32 + segbits : 31;
  int in_b = 0;
  rc = lg_radix_sort(ws.rs, M, bits, st, launches, &in_b);
  if (rc) return rc;
  const unsigned long long* keys = in_b ? ws.rs.keysB.as<unsigned long long>() : ws.rs.keysA.as<unsigned long long>();
  const unsigned int* vals = in_b ? ws.rs.valsB.as<unsigned int>() : ws.rs.valsA.as<unsigned int>();
  int ncb = lg_div_up(M, VC_TILE);
  LG_CHECK(ws.block_sums.ensure((size_t)ncb * sizeof(unsigned int), st));
  vb_count_kernel<<<ncb, VC_NT, 0, st>>>(keys, M, ws.block_sums.as<unsigned int>());
  rs_scan_kernel<<<1, 1024, 0, st>>>(ws.block_sums.as<unsigned int>(), ncb);
  vb_centroid_kernel<<<ncb, VC_NT, 0, st>>>(d_in, keys, vals, M, ws.block_sums.as<unsigned int>(), d_seg_off, d_out, d_out_start, d_out_end);
  (*launches) += 3;
  LG_CHECK(cudaGetLastError());
  return LOAM_OK;
}


int lg_radix_sort_shifts(RadixWs& ws, int n, const int* shifts, int nshifts, cudaStream_t st, long long* launches, int* result_in_b) {
  if (n <= RSS_CAP && nshifts <= 8) return rs_small_launch(ws, n, shifts, nshifts, st, launches, result_in_b);
  int nblocks = lg_div_up(n, RS_TILE);
  LG_CHECK(ws.hist.ensure((size_t)256 * nblocks * sizeof(unsigned int), st));
  unsigned long long* ka = ws.keysA.as<unsigned long long>();
  unsigned long long* kb = ws.keysB.as<unsigned long long>();
  unsigned int* va = ws.valsA.as<unsigned int>();
  unsigned int* vb = ws.valsB.as<unsigned int>();
  int flip = 0;
  for (int p = 0; p < nshifts; p++) {
    rs_hist_kernel<<<nblocks, RS_NT, 0, st>>>(ka, n, shifts[p], ws.hist.as<unsigned int>(), nblocks);
    rs_scan_kernel<<<1, 1024, 0, st>>>(ws.hist.as<unsigned int>(), 256 * nblocks);
    rs_scatter_kernel<<<nblocks, RS_NT, 0, st>>>(ka, va, kb, vb, n, shifts[p], ws.hist.as<unsigned int>(), nblocks);
    (*launches) += 3;
    std::swap(ka, kb);
    std::swap(va, vb);
    flip ^= 1;
  }
  LG_CHECK(cudaGetLastError());
  *result_in_b = flip;
  return LOAM_OK;
}

// d_in = [old points of all segments (segment-major) | new points of all segments]; d_flags[0] != 0 afterwards means
// the fast path does not apply (caller re-runs lg_vox_big on the same input).  d_seg_off = merged offsets.
int lg_vox_merge(VoxBigWs& ws, const float4* d_in, const int* d_seg_off_old, const int* d_seg_off_new, const int* d_seg_off,
                 const float* d_seg_leaf, int nseg, int n_old, int n_new, float4* d_out, int* d_out_start, int* d_out_end, int* d_flags,
                 cudaStream_t st, long long* launches) {
  const int M = n_old + n_new;
  if (nseg <= 0) return LOAM_OK;
  LgProfScope prof_scope(LGK_VOXEL, st, (double)M);
  LG_CHECK(ws.bb.ensure((size_t)nseg * 6 * sizeof(int), st));
  LG_CHECK(cudaMemsetAsync(d_flags, 0, 4, st));
  vb_init_kernel<<<lg_div_up(nseg * 6, 256), 256, 0, st>>>(ws.bb.as<int>(), nseg, d_out_start, d_out_end);
  (*launches)++;
  if (M <= 0) return LOAM_OK;
  int rc = lg_radix_ensure(ws.rs, std::max(n_new, 1), st);
  if (rc) return rc;
  LG_CHECK(ws.keys_old.ensure((size_t)(n_old + 1) * 8, st));
  LG_CHECK(ws.vals_old.ensure((size_t)(n_old + 1) * 4, st));
  LG_CHECK(ws.keys_m.ensure((size_t)(M + 1) * 8, st));
  LG_CHECK(ws.vals_m.ensure((size_t)(M + 1) * 4, st));
  if (n_old > 0) vb_bbox_kernel<<<lg_div_up(n_old, 256), 256, 0, st>>>(d_in, d_seg_off_old, nseg, n_old, ws.bb.as<int>());
  if (n_new > 0) vb_bbox_kernel<<<lg_div_up(n_new, 256), 256, 0, st>>>(d_in + n_old, d_seg_off_new, nseg, n_new, ws.bb.as<int>());
  if (n_old > 0)
    vm_key_kernel<<<lg_div_up(n_old, 256), 256, 0, st>>>(d_in, d_seg_off_old, d_seg_leaf, nseg, n_old, 0, ws.bb.as<int>(),
                                                         ws.keys_old.as<unsigned long long>(), ws.vals_old.as<unsigned int>(), 1, d_flags);
  if (n_new > 0)
    vm_key_kernel<<<lg_div_up(n_new, 256), 256, 0, st>>>(d_in + n_old, d_seg_off_new, d_seg_leaf, nseg, n_new, n_old, ws.bb.as<int>(),
                                                         ws.rs.keysA.as<unsigned long long>(), ws.rs.valsA.as<unsigned int>(), 0, d_flags);
  (*launches) += 4;
  int in_b = 0;
  if (n_new > 1) {
    int segbits = 0;
    while ((1 << segbits) < nseg) segbits++;
    int shifts[8], ns = 0;
    for (int s = 0; s < 24; s += 8) shifts[ns++] = s;
    for (int s = 0; s < segbits; s += 8) shifts[ns++] = 32 + s;
    rc = lg_radix_sort_shifts(ws.rs, n_new, shifts, ns, st, launches, &in_b);
    if (rc) return rc;
  }
  const unsigned long long* kn = in_b ? ws.rs.keysB.as<unsigned long long>() : ws.rs.keysA.as<unsigned long long>();
  const unsigned int* vn = in_b ? ws.rs.valsB.as<unsigned int>() : ws.rs.valsA.as<unsigned int>();
  vm_merge_kernel<<<lg_div_up(M, 256), 256, 0, st>>>(ws.keys_old.as<unsigned long long>(), ws.vals_old.as<unsigned int>(), n_old, kn, vn, n_new,
                                                     ws.keys_m.as<unsigned long long>(), ws.vals_m.as<unsigned int>());
  int ncb = lg_div_up(M, VC_TILE);
  LG_CHECK(ws.block_sums.ensure((size_t)ncb * sizeof(unsigned int), st));
  vb_count_kernel<<<ncb, VC_NT, 0, st>>>(ws.keys_m.as<unsigned long long>(), M, ws.block_sums.as<unsigned int>());
  rs_scan_kernel<<<1, 1024, 0, st>>>(ws.block_sums.as<unsigned int>(), ncb);
  vb_centroid_kernel<<<ncb, VC_NT, 0, st>>>(d_in, ws.keys_m.as<unsigned long long>(), ws.vals_m.as<unsigned int>(), M,
                                            ws.block_sums.as<unsigned int>(), d_seg_off, d_out, d_out_start, d_out_end);
  (*launches) += 4;
  LG_CHECK(cudaGetLastError());
  return LOAM_OK;
}

int lg_gather(const CopyEnt* d_ents, int nent, int max_n, float4* d_dst, cudaStream_t st, long long* launches) {
  if (nent <= 0 || max_n <= 0) return LOAM_OK;
  LgProfScope prof_scope(LGK_GATHER, st, 0.0);
  dim3 grid(std::max(1, std::min(lg_div_up(max_n, 256), 64)), std::min(nent, 1024));
  gather_kernel<<<grid, 256, 0, st>>>(d_ents, nent, d_dst);
  (*launches)++;
  LG_CHECK(cudaGetLastError());
  return LOAM_OK;
}
