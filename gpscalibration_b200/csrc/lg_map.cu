// K6 / K7 / K10 / K11 — scan-to-map kernels, the B200 replacement for the hot loops of laserMapping.cpp:
//   map_stack_kernel     LM:467-477,726-734  pointAssociateToMap followed by pointAssociateTobeMapped (the round trip is
//                                            kept so voxel membership matches the reference to the ulp, Appendix B.11)
//   grid_*_kernel        LM:750-751          spatial index over the gathered local map (replaces KdTreeFLANN::
//                                            setInputCloud): 1 m voxel hash (open addressing) + cell-sorted float4 copy
//                                            carrying the original index in .w — index build traffic 36 T bytes
//   map_knn_kernel       LM:760,867          exact 5-NN, one warp per query: each lane probes one of the 27 neighbour
//                                            cells, keeps a sorted top-5 of (d2, index) keys, warp-merge by shuffles.
//                                            Only points with d2 < 1 m^2 can take part in an ACCEPTED correspondence
//                                            (LM:762,869), so the 27-cell search is exact for everything the
//                                            reference uses; rejected queries report -1.
//   map_fit_kernel       LM:763-964          3x3 covariance + Jacobi eigen line fit / 5x3 Householder plane fit,
//                                            weights, Jacobian row, 21 + 6 term reduction
//   map_insert_kernel    LM:1023-1059        pointAssociateToMap + cube index of every stack point
//   map_register_kernel  LM:1103-1106        full-resolution cloud into the map frame
#include "lg_linalg.cuh"
#include "lg_map.h"
#include "lg_reduce.cuh"

namespace {

// LM:244-262 with the six sin/cos values evaluated once on the host (libm), as they do not depend on the point.
__device__ __forceinline__ float4 assoc_to_map(const MapT& T, float4 pi) {
  float x1 = T.sc.crz * pi.x - T.sc.srz * pi.y;
  float y1 = T.sc.srz * pi.x + T.sc.crz * pi.y;
  float z1 = pi.z;
  float x2 = x1;
  float y2 = T.sc.crx * y1 - T.sc.srx * z1;
  float z2 = T.sc.srx * y1 + T.sc.crx * z1;
  float4 po;
  po.x = T.sc.cry * x2 + T.sc.sry * z2 + T.t[3];
  po.y = y2 + T.t[4];
  po.z = -T.sc.sry * x2 + T.sc.cry * z2 + T.t[5];
  po.w = pi.w;
  return po;
}
// LM:264-282
__device__ __forceinline__ float4 assoc_tobe_mapped(const MapT& T, float4 pi) {
  float x1 = T.sc.cry * (pi.x - T.t[3]) - T.sc.sry * (pi.z - T.t[5]);
  float y1 = pi.y - T.t[4];
  float z1 = T.sc.sry * (pi.x - T.t[3]) + T.sc.cry * (pi.z - T.t[5]);
  float x2 = x1;
  float y2 = T.sc.crx * y1 + T.sc.srx * z1;
  float z2 = -T.sc.srx * y1 + T.sc.crx * z1;
  float4 po;
  po.x = T.sc.crz * x2 + T.sc.srz * y2;
  po.y = -T.sc.srz * x2 + T.sc.crz * y2;
  po.z = z2;
  po.w = pi.w;
  return po;
}

__global__ void map_stack_kernel(MapT T, const float4* __restrict__ in0, float4* __restrict__ out0, int n0, const float4* __restrict__ in1,
                                 float4* __restrict__ out1, int n1) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n0) {
    out0[i] = assoc_tobe_mapped(T, assoc_to_map(T, in0[i]));
  } else if (i < n0 + n1) {
    i -= n0;
    out1[i] = assoc_tobe_mapped(T, assoc_to_map(T, in1[i]));
  }
}

__global__ void map_register_kernel(MapT T, const float4* __restrict__ in, float4* __restrict__ out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = assoc_to_map(T, in[i]);
}

// ---------------------------------------------------------------------------------------------- voxel hash grid
constexpr unsigned long long EMPTY = ~0ull;
constexpr int COFF = 1 << 20;

__device__ __forceinline__ unsigned long long cell_key(int ix, int iy, int iz) {
  return ((unsigned long long)(unsigned int)(ix + COFF) << 42) | ((unsigned long long)(unsigned int)(iy + COFF) << 21) |
         (unsigned long long)(unsigned int)(iz + COFF);
}
__device__ __forceinline__ unsigned int cell_hash(unsigned long long k, int bits) { return (unsigned int)((k * 0x9E3779B97F4A7C15ull) >> (64 - bits)); }

// The corner grid and the surf grid are always built together: every build kernel serves both (CTAs below `split` work
// on grid 0, the rest on grid 1), so a map refresh costs four launches instead of ten.
struct GridJob {
  GridD g[2];
  const float4* pts[2];
  int n[2];
};
__global__ void grid_init_kernel(GridJob J, int split) {
  const int w = (int)blockIdx.x >= split;
  const GridD& g = J.g[w];
  const int s = ((int)blockIdx.x - (w ? split : 0)) * blockDim.x + threadIdx.x;  // the slot count is a multiple of the block size
  uint4* raw = reinterpret_cast<uint4*>(&g.slots[s]);
  raw[0] = make_uint4(0xffffffffu, 0xffffffffu, 0u, 0xffffffffu);  // key = EMPTY, count = 0, line = -1
  raw[1] = make_uint4(0u, 0u, 0u, 0u);
  g.fill[s] = 0;
  if (s == 0) g.cursor[0] = g.cursor[1] = 0;
}
__global__ void grid_count_kernel(GridJob J, int split) {
  const int w = (int)blockIdx.x >= split;
  const GridD& g = J.g[w];
  const int i = ((int)blockIdx.x - (w ? split : 0)) * blockDim.x + threadIdx.x;
  if (i >= J.n[w]) return;
  float4 p = J.pts[w][i];
  unsigned long long key = cell_key((int)floorf(p.x), (int)floorf(p.y), (int)floorf(p.z));
  unsigned int h = cell_hash(key, g.bits);
  const unsigned int mask = (1u << g.bits) - 1u;
  while (true) {
    unsigned long long prev = atomicCAS(&g.slots[h].key, EMPTY, key);
    if (prev == EMPTY || prev == key) break;
    h = (h + 1) & mask;
  }
  atomicAdd(&g.slots[h].count, 1);
  g.slot_of[i] = (int)h;
}
__global__ void grid_alloc_kernel(GridJob J, int split) {
  const int w = (int)blockIdx.x >= split;
  const GridD& g = J.g[w];
  const int s = ((int)blockIdx.x - (w ? split : 0)) * blockDim.x + threadIdx.x;
  const int c = g.slots[s].count;
  if (c > 1) {
    const int line = atomicAdd(&g.cursor[1], 1);
    g.slots[s].line = line;
    if (c > GRID_INLINE) g.ovf_start[line] = atomicAdd(&g.cursor[0], c);
  }
  const unsigned int b = __ballot_sync(0xffffffffu, c > 0);
  if ((threadIdx.x & 31) == 0) const_cast<unsigned int*>(g.occ)[s >> 5] = b;
}
__global__ void grid_fill_kernel(GridJob J, int split) {
  const int w = (int)blockIdx.x >= split;
  const GridD& g = J.g[w];
  const int i = ((int)blockIdx.x - (w ? split : 0)) * blockDim.x + threadIdx.x;
  if (i >= J.n[w]) return;
  int s = g.slot_of[i];
  int pos = atomicAdd(&g.fill[s], 1);
  float4 p = J.pts[w][i];
  float4 e = make_float4(p.x, p.y, p.z, __int_as_float(i));
  const int cnt = g.slots[s].count, line = g.slots[s].line;
  if (pos == 0) g.slots[s].p0 = e;
  else if (pos < GRID_INLINE) g.lines[line].pts[pos - 1] = e;
  if (cnt > GRID_INLINE) g.sorted[g.ovf_start[line] + pos] = e;
}

// ---------------------------------------------------------------------------------------------- exact 5-NN
constexpr int KNN_WARPS = 8;

// Exact 5-NN, EIGHT lanes per query (four queries per warp), two phases.  Probe: every lane resolves 3-4 of the 27
// neighbour cells -- one L2-resident occupancy bit decides whether the cell exists at all (most do not), one 32-byte
// sector {key, count, line, point 0} resolves it.  Scan: the cells that hold more than one point are then walked by
// the whole group TOGETHER, lane i taking point i of the cell's 128-byte line (one coalesced read, and every lane sees
// the same number of candidates whatever the occupancy of "its" cells).  Each lane keeps a sorted top-5 of (d2, index)
// keys in registers (branch-free min/max insertion); the eight lanes are merged with masked warp reductions.
// Variants measured and dropped on a 20 M-point map: a full warp per query (same time, 3x the instructions), all
// probes of a lane issued up front (more registers, slower), a block-local hash layout (slower build, no gain).
template <int KNN_SUB>  // lanes per query: 8 (four queries per warp) for large batches, 16 when the batch cannot fill the GPU
__global__ void __launch_bounds__(KNN_WARPS * 32) map_knn_kernel(MapT T, const float4* __restrict__ corner_stack, int n_cs,
                                                                  const float4* __restrict__ surf_stack, int n_ss, GridD gc, GridD gs,
                                                                  int* __restrict__ nbr /* [n_cs + n_ss][5] */) {
  constexpr int KNN_QPW = 32 / KNN_SUB;  // queries per warp
  constexpr unsigned int SUBMASK = KNN_SUB == 32 ? 0xffffffffu : ((1u << (KNN_SUB & 31)) - 1u);
  const int lane = threadIdx.x & 31, sub = lane & (KNN_SUB - 1), grp = lane / KNN_SUB;
  const int q = (blockIdx.x * KNN_WARPS + (threadIdx.x >> 5)) * KNN_QPW + grp;
  const unsigned int gmask = SUBMASK << (grp * (KNN_SUB & 31));
  const bool active = q < n_cs + n_ss;
  const bool is_c = q < n_cs;
  const GridD& g = is_c ? gc : gs;
  float4 sel = make_float4(0.f, 0.f, 0.f, 0.f);
  if (active) sel = assoc_to_map(T, is_c ? corner_stack[q] : surf_stack[q - n_cs]);
  unsigned long long k0 = EMPTY, k1 = EMPTY, k2 = EMPTY, k3 = EMPTY, k4 = EMPTY;
  auto offer = [&](float4 p) {
    const float d2 = lg_sqdist(p.x, p.y, p.z, sel.x, sel.y, sel.z);
    if (d2 < 1.0f) {
      const unsigned long long c = lg_pack_nbr(d2, __float_as_int(p.w));
      if (c < k4) {  // sorted insertion without branches: new_i = min(max(c, k_{i-1}), k_i)
        const unsigned long long n4 = min(max(c, k3), k4), n3 = min(max(c, k2), k3), n2 = min(max(c, k1), k2), n1 = min(max(c, k0), k1);
        k0 = min(c, k0);
        k1 = n1; k2 = n2; k3 = n3; k4 = n4;
      }
    }
  };
  constexpr int CPL = (27 + KNN_SUB - 1) / KNN_SUB;  // cells per lane
  int ccnt[CPL], cline[CPL];
  {
    const int bx = (int)floorf(sel.x), by = (int)floorf(sel.y), bz = (int)floorf(sel.z);
    const unsigned int mask = (1u << g.bits) - 1u;
#pragma unroll
    for (int j = 0; j < CPL; j++) {
      const int c = sub + j * KNN_SUB;
      ccnt[j] = 0;
      cline[j] = -1;
      if (!active || c >= 27) continue;
      const unsigned long long key = cell_key(bx + (c % 3) - 1, by + ((c / 3) % 3) - 1, bz + (c / 9) - 1);
      unsigned int h = cell_hash(key, g.bits);
      if (!((__ldg(&g.occ[h >> 5]) >> (h & 31)) & 1u)) continue;  // home slot empty => cell absent (linear probing)
      uint4 hd;
      float4 p0;
      unsigned long long kk;
      while (true) {
        // one 32-byte sector: {key, count, line} + point 0, two independent 16-byte loads
        hd = __ldg(reinterpret_cast<const uint4*>(&g.slots[h]));
        p0 = __ldg(&g.slots[h].p0);
        kk = ((unsigned long long)hd.y << 32) | hd.x;
        if (kk == key || kk == EMPTY) break;
        h = (h + 1) & mask;
      }
      if (kk != key) continue;
      offer(p0);
      ccnt[j] = (int)hd.z;
      cline[j] = (int)hd.w;
    }
  }
#pragma unroll
  for (int j = 0; j < CPL; j++) {
    unsigned int todo = (__ballot_sync(0xffffffffu, ccnt[j] > 1) >> (grp * (KNN_SUB & 31))) & SUBMASK;
    while (todo) {  // uniform inside the 8-lane group
      const int src = grp * KNN_SUB + __ffs(todo) - 1;
      todo &= todo - 1;
      const int cnt = __shfl_sync(gmask, ccnt[j], src);
      const int line = __shfl_sync(gmask, cline[j], src);
      if (sub < min(cnt - 1, GRID_LINE)) offer(__ldg(&g.lines[line].pts[sub]));
      if (cnt > GRID_INLINE) {  // crowded cell: the rest comes from the cell-sorted array, eight points per step
        const int start = __ldg(&g.ovf_start[line]);
        for (int i = GRID_INLINE + sub; i < cnt; i += KNN_SUB) offer(__ldg(&g.sorted[start + i]));
      }
    }
  }
  // merge inside the 8-lane group: five rounds of "minimum of the lanes' heads"; keys are unique (they embed the index)
  unsigned long long res[5];
#pragma unroll
  for (int r = 0; r < 5; r++) {
    const unsigned int hi = (unsigned int)(k0 >> 32);
    const unsigned int mhi = __reduce_min_sync(gmask, hi);
    const unsigned int lo = (hi == mhi) ? (unsigned int)k0 : 0xffffffffu;
    const unsigned int mlo = __reduce_min_sync(gmask, lo);
    const unsigned long long m = ((unsigned long long)mhi << 32) | mlo;
    res[r] = m;
    if (k0 == m && m != EMPTY) {
      k0 = k1; k1 = k2; k2 = k3; k3 = k4; k4 = EMPTY;
    }
  }
  if (sub == 0 && active) {
    bool ok = res[4] != EMPTY;
#pragma unroll
    for (int r = 0; r < 5; r++) nbr[(size_t)q * 5 + r] = ok ? lg_nbr_idx(res[r]) : -1;
  }
}

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

constexpr int FIT_NT = 128;

__global__ void __launch_bounds__(FIT_NT) map_fit_kernel(MapT T, const float4* __restrict__ corner_stack, int n_cs,
                                                          const float4* __restrict__ surf_stack, int n_ss, const float4* __restrict__ corner_map,
                                                          const float4* __restrict__ surf_map, const int* __restrict__ nbr,
                                                          double* __restrict__ partials, unsigned int* __restrict__ ticket, double* __restrict__ out28,
                                                          unsigned long long seq, PeerXchg px) {
  Acc28 acc;
  acc.clear();
  // grid-stride over the queries: the grid is capped (a few CTAs per SM) so the last-CTA reduction stays short
  for (int q = blockIdx.x * FIT_NT + threadIdx.x; q < n_cs + n_ss; q += gridDim.x * FIT_NT) {
  float4 ori, coef;
  bool keep = false;
  if (nbr[(size_t)q * 5] >= 0) {
    const bool is_c = q < n_cs;
    ori = is_c ? corner_stack[q] : surf_stack[q - n_cs];
    const float4 sel = assoc_to_map(T, ori);
    const float4* mp = is_c ? corner_map : surf_map;
    float px[5], py[5], pz[5];
#pragma unroll
    for (int j = 0; j < 5; j++) {
      float4 p = mp[nbr[(size_t)q * 5 + j]];
      px[j] = p.x; py[j] = p.y; pz[j] = p.z;
    }
    if (is_c) {  // LM:763-861
      float cx = 0, cy = 0, cz = 0;
#pragma unroll
      for (int j = 0; j < 5; j++) {
        cx += px[j]; cy += py[j]; cz += pz[j];
      }
      cx /= 5; cy /= 5; cz /= 5;
      float a11 = 0, a12 = 0, a13 = 0, a22 = 0, a23 = 0, a33 = 0;
#pragma unroll
      for (int j = 0; j < 5; j++) {
        float ax = px[j] - cx, ay = py[j] - cy, az = pz[j] - cz;
        a11 += ax * ax; a12 += ax * ay; a13 += ax * az;
        a22 += ay * ay; a23 += ay * az; a33 += az * az;
      }
      a11 /= 5; a12 /= 5; a13 /= 5; a22 /= 5; a23 /= 5; a33 /= 5;
      float A1[9] = {a11, a12, a13, a12, a22, a23, a13, a23, a33};
      float D1[3], V1[9];
      lg_jacobi_eigen<3>(A1, D1, V1);
      if (D1[0] > 3 * D1[1]) {
        float x1 = (float)(cx + 0.1 * V1[0]), y1 = (float)(cy + 0.1 * V1[1]), z1 = (float)(cz + 0.1 * V1[2]);
        float x2 = (float)(cx - 0.1 * V1[0]), y2 = (float)(cy - 0.1 * V1[1]), z2 = (float)(cz - 0.1 * V1[2]);
        float la, lb, lc, ld2;
        line_coeff(sel.x, sel.y, sel.z, x1, y1, z1, x2, y2, z2, la, lb, lc, ld2);
        float s = (float)(1 - 0.9 * fabsf(ld2));
        coef = make_float4(s * la, s * lb, s * lc, s * ld2);
        keep = s > 0.1;
      }
    } else {  // LM:870-919
      float A0[15], B0[5] = {-1, -1, -1, -1, -1}, X0[3];
#pragma unroll
      for (int j = 0; j < 5; j++) {
        A0[j * 3 + 0] = px[j]; A0[j * 3 + 1] = py[j]; A0[j * 3 + 2] = pz[j];
      }
      lg_qr_solve<5, 3>(A0, B0, X0);
      float pa = X0[0], pb = X0[1], pc = X0[2], pd = 1;
      float ps = sqrtf(pa * pa + pb * pb + pc * pc);
      pa /= ps; pb /= ps; pc /= ps; pd /= ps;
      bool planeValid = true;
#pragma unroll
      for (int j = 0; j < 5; j++)
        if (fabsf(pa * px[j] + pb * py[j] + pc * pz[j] + pd) > 0.2) planeValid = false;
      if (planeValid) {
        float pd2 = pa * sel.x + pb * sel.y + pc * sel.z + pd;
        float s = (float)(1 - 0.9 * fabsf(pd2) / sqrtf(sqrtf(sel.x * sel.x + sel.y * sel.y + sel.z * sel.z)));
        coef = make_float4(s * pa, s * pb, s * pc, s * pd2);
        keep = s > 0.1;
      }
    }
  }
  if (keep) {  // LM:940-964
    const float srx = T.sc.srx, crx = T.sc.crx, sry = T.sc.sry, cry = T.sc.cry, srz = T.sc.srz, crz = T.sc.crz;
    const float4 p = ori, c = coef;
    float a[6];
    a[0] = (crx * sry * srz * p.x + crx * crz * sry * p.y - srx * sry * p.z) * c.x + (-srx * srz * p.x - crz * srx * p.y - crx * p.z) * c.y +
           (crx * cry * srz * p.x + crx * cry * crz * p.y - cry * srx * p.z) * c.z;
    a[1] = ((cry * srx * srz - crz * sry) * p.x + (sry * srz + cry * crz * srx) * p.y + crx * cry * p.z) * c.x +
           ((-cry * crz - srx * sry * srz) * p.x + (cry * srz - crz * srx * sry) * p.y - crx * sry * p.z) * c.z;
    a[2] = ((crz * srx * sry - cry * srz) * p.x + (-cry * crz - srx * sry * srz) * p.y) * c.x + (crx * crz * p.x - crx * srz * p.y) * c.y +
           ((sry * srz + cry * crz * srx) * p.x + (crz * sry - cry * srx * srz) * p.y) * c.z;
    a[3] = c.x;
    a[4] = c.y;
    a[5] = c.z;
    acc.add_row(a, -c.w);
  }
  }
  lg_reduce28<FIT_NT>(acc, partials, ticket, out28, seq, &px);
}

// LM:1023-1059: map-frame point and cube index; key 0xFFFFFFFF.. sorts dropped points to the end.
__global__ void map_insert_kernel(MapT T, CubeGeom cg, const float4* __restrict__ corner_stack, int n_cs, const float4* __restrict__ surf_stack,
                                  int n_ss, float4* __restrict__ sel_out, unsigned long long* __restrict__ keys, unsigned int* __restrict__ vals) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_cs + n_ss) return;
  const bool is_c = i < n_cs;
  float4 sel = assoc_to_map(T, is_c ? corner_stack[i] : surf_stack[i - n_cs]);
  int cubeI = int((sel.x + 25.0) / 50.0) + cg.cenW;
  int cubeJ = int((sel.y + 25.0) / 50.0) + cg.cenH;
  int cubeK = int((sel.z + 25.0) / 50.0) + cg.cenD;
  if (sel.x + 25.0 < 0) cubeI--;
  if (sel.y + 25.0 < 0) cubeJ--;
  if (sel.z + 25.0 < 0) cubeK--;
  unsigned long long key = 0x3fffull;  // dropped
  if (cubeI >= 0 && cubeI < cg.W && cubeJ >= 0 && cubeJ < cg.H && cubeK >= 0 && cubeK < cg.D)
    key = (unsigned long long)(cubeI + cg.W * cubeJ + cg.W * cg.H * cubeK);
  // corner points sort before surf points: bit 14 = type, bits 0..13 = cube (4851 < 16383)
  keys[i] = key | (is_c ? 0ull : (1ull << 14));
  vals[i] = (unsigned int)i;
  sel_out[i] = sel;
}

// After the sort: heads of equal-key runs -> (key, start) records; the host derives the run lengths.
__global__ void map_runs_kernel(const unsigned long long* __restrict__ keys, const unsigned int* __restrict__ vals, const float4* __restrict__ sel,
                                int n, float4* __restrict__ sorted_sel, int* __restrict__ n_runs, int2* __restrict__ runs, int cap_runs) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  sorted_sel[i] = sel[vals[i]];
  if (i == 0 || keys[i] != keys[i - 1]) {
    int r = atomicAdd(n_runs, 1);
    if (r < cap_runs) runs[r] = make_int2((int)keys[i], i);
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------------------ host side
int lg_map_stack_launch(const MapT& T, const float4* in0, float4* out0, int n0, const float4* in1, float4* out1, int n1, cudaStream_t st,
                        long long* launches) {
  if (n0 + n1 <= 0) return LOAM_OK;
  LgProfScope prof_scope(LGK_MAP_STACK, st, (double)(n0 + n1));
  map_stack_kernel<<<lg_div_up(n0 + n1, 256), 256, 0, st>>>(T, in0, out0, n0, in1, out1, n1);
  (*launches)++;
  LG_CHECK(cudaGetLastError());
  return LOAM_OK;
}

int lg_map_register_launch(const MapT& T, const float4* in, float4* out, int n, cudaStream_t st, long long* launches) {
  if (n <= 0) return LOAM_OK;
  LgProfScope prof_scope(LGK_MAP_STACK, st, (double)n);
  map_register_kernel<<<lg_div_up(n, 256), 256, 0, st>>>(T, in, out, n);
  (*launches)++;
  LG_CHECK(cudaGetLastError());
  return LOAM_OK;
}

static int grid_prepare(GridWs& ws, int n, cudaStream_t st, size_t* slots_out) {
  int bits = 10;
  while ((1 << bits) < 2 * n) bits++;
  const size_t slots = (size_t)1 << bits;
  LG_CHECK(ws.keys.ensure(slots * sizeof(GridSlot), st));
  LG_CHECK(ws.lines.ensure((size_t)(n + 1) * sizeof(GridLine), st));
  LG_CHECK(ws.ovf.ensure((size_t)(n + 1) * 4, st));
  LG_CHECK(ws.ints.ensure((slots + 8 + slots / 32) * 4, st));
  LG_CHECK(ws.slot_of.ensure((size_t)(n + 1) * 4, st));
  LG_CHECK(ws.sorted.ensure((size_t)(n + 1) * 16, st));
  GridD& g = ws.d;
  g.slots = ws.keys.as<GridSlot>();
  g.lines = ws.lines.as<GridLine>();
  g.ovf_start = ws.ovf.as<int>();
  g.fill = ws.ints.as<int>();
  g.cursor = g.fill + slots;
  g.occ = reinterpret_cast<const unsigned int*>(g.cursor + 8);
  g.slot_of = ws.slot_of.as<int>();
  g.sorted = ws.sorted.as<float4>();
  g.bits = bits;
  g.n = n;
  *slots_out = slots;
  return LOAM_OK;
}

int lg_grid_reserve(GridWs& ws, int n, cudaStream_t st) {
  size_t slots = 0;
  return grid_prepare(ws, n, st, &slots);
}

int lg_grid_build2(GridWs& ws0, const float4* pts0, int n0, GridWs& ws1, const float4* pts1, int n1, cudaStream_t st, long long* launches) {
  size_t slots0 = 0, slots1 = 0;
  int rc = grid_prepare(ws0, n0, st, &slots0);
  if (rc) return rc;
  rc = grid_prepare(ws1, n1, st, &slots1);
  if (rc) return rc;
  GridJob J;
  J.g[0] = ws0.d; J.g[1] = ws1.d;
  J.pts[0] = pts0; J.pts[1] = pts1;
  J.n[0] = n0; J.n[1] = n1;
  LgProfScope prof_scope(LGK_GRID, st, (double)(n0 + n1));
  const int sb0 = (int)(slots0 / 256), sb1 = (int)(slots1 / 256);
  const int pb0 = lg_div_up(n0, 256), pb1 = lg_div_up(n1, 256);
  grid_init_kernel<<<sb0 + sb1, 256, 0, st>>>(J, sb0);
  (*launches)++;
  if (pb0 + pb1 > 0) {
    grid_count_kernel<<<pb0 + pb1, 256, 0, st>>>(J, pb0);
    grid_alloc_kernel<<<sb0 + sb1, 256, 0, st>>>(J, sb0);
    grid_fill_kernel<<<pb0 + pb1, 256, 0, st>>>(J, pb0);
    (*launches) += 3;
  }
  LG_CHECK(cudaGetLastError());
  return LOAM_OK;
}

int lg_map_iter_launch(MapIterWs& ws, const MapT& T, const float4* corner_stack, int n_cs, const float4* surf_stack, int n_ss, const GridD& gc,
                       const GridD& gs, const float4* corner_map, const float4* surf_map, double* out28, unsigned long long seq, cudaStream_t st,
                       long long* launches, const PeerXchg* px) {
  const int nq = n_cs + n_ss;
  const int nb = std::max(1, std::min(lg_div_up(nq, FIT_NT), 148 * 8));
  LG_CHECK(ws.nbr.ensure((size_t)(nq + 1) * 5 * 4, st));
  LG_CHECK(ws.partials.ensure((size_t)nb * 28 * 8, st));
  if (!ws.ticket.p) {
    LG_CHECK(ws.ticket.ensure(4, st));
    LG_CHECK(cudaMemsetAsync(ws.ticket.p, 0, 4, st));
  }
  if (nq > 0) {
    LgProfScope prof_scope(LGK_MAP_KNN, st, (double)nq);
    if (nq <= 148 * 128)  // fewer than ~four resident warps per scheduler at 8 lanes per query: spread every query wider
      map_knn_kernel<16><<<lg_div_up(nq, KNN_WARPS * 2), KNN_WARPS * 32, 0, st>>>(T, corner_stack, n_cs, surf_stack, n_ss, gc, gs, ws.nbr.as<int>());
    else
      map_knn_kernel<8><<<lg_div_up(nq, KNN_WARPS * 4), KNN_WARPS * 32, 0, st>>>(T, corner_stack, n_cs, surf_stack, n_ss, gc, gs, ws.nbr.as<int>());
    (*launches)++;
  }
  LgProfScope prof_scope(LGK_MAP_FIT, st, (double)nq);
  PeerXchg none;
  memset(&none, 0, sizeof(none));
  map_fit_kernel<<<nb, FIT_NT, 0, st>>>(T, corner_stack, n_cs, surf_stack, n_ss, corner_map, surf_map, ws.nbr.as<int>(), ws.partials.as<double>(),
                                        ws.ticket.as<unsigned int>(), out28, seq, px ? *px : none);
  (*launches)++;
  LG_CHECK(cudaGetLastError());
  return LOAM_OK;
}

int lg_map_insert_launch(const MapT& T, const CubeGeom& cg, const float4* corner_stack, int n_cs, const float4* surf_stack, int n_ss,
                         float4* sel_out, unsigned long long* keys, unsigned int* vals, cudaStream_t st, long long* launches) {
  const int n = n_cs + n_ss;
  if (n <= 0) return LOAM_OK;
  LgProfScope prof_scope(LGK_INSERT, st, (double)n);
  map_insert_kernel<<<lg_div_up(n, 256), 256, 0, st>>>(T, cg, corner_stack, n_cs, surf_stack, n_ss, sel_out, keys, vals);
  (*launches)++;
  LG_CHECK(cudaGetLastError());
  return LOAM_OK;
}

int lg_map_runs_launch(const unsigned long long* keys, const unsigned int* vals, const float4* sel, int n, float4* sorted_sel, int* n_runs,
                       int2* runs, int cap_runs, cudaStream_t st, long long* launches) {
  LG_CHECK(cudaMemsetAsync(n_runs, 0, 4, st));
  if (n <= 0) return LOAM_OK;
  LgProfScope prof_scope(LGK_INSERT, st, 0.0);
  map_runs_kernel<<<lg_div_up(n, 256), 256, 0, st>>>(keys, vals, sel, n, sorted_sel, n_runs, runs, cap_runs);
  (*launches)++;
  LG_CHECK(cudaGetLastError());
  return LOAM_OK;
}
