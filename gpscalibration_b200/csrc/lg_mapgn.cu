// K7 / K10 — the scan-to-map Gauss-Newton loop of laserMapping.cpp (LM:749-1020) as ONE persistent kernel over a
// cell-sorted (CSR) grid of the local map:
//
//   csr_*_kernel     LM:750-751  replaces KdTreeFLANN::setInputCloud.  The corner map and the surf map are bucketed into
//                                1 m cells of a dense grid over their bounding boxes (x fastest, then y, then z): a
//                                histogram, one exclusive scan over BOTH tables, a scatter.  Output: the points in cell
//                                order (float4 {x, y, z, original index}) and, per cell, the END of its run.  The 27 cells
//                                around a query are then nine contiguous x-runs, each found with two loads of the table
//                                (E[row + cx] and E[row + cx + 3]) -- no hashing, no probing, no per-cell indirection, and
//                                queries that are neighbours in space read neighbouring table entries and runs.
//                                Index build traffic 36 T bytes (16 T read, 4 T cell ids, 16 T reordered write).
//   map_gn_kernel    LM:753-1017 every iteration: pointAssociateToMap of the stack points (LM:756, 865), exact 5-NN
//                                (LM:760, 867), line / plane fit (LM:763-919) and Jacobian row (LM:940-964), the
//                                21 + 6 + 1 sums (LM:965-967) warp -> CTA -> grid, then -- on
//                                the device, by the last CTA to arrive -- the 6x6 solve, degeneracy projection, pose
//                                update and convergence test (LM:968-1017), the six sin / cos of the new pose, and a grid
//                                barrier into the next iteration.  The host sees one mailbox write per mapping run.
//                                With the map sharded over several GPUs every rank keeps the whole stack,
//                                evaluates the queries whose map-frame x falls into its slab (re-routed every iteration
//                                with the same arithmetic), and the last CTA exchanges the 28 sums with its peers through
//                                NVLink peer memory before the solve: reduction + collective + solve in one kernel.
//                                Two layouts of the search.  Up to 2^19 stack points: eight lanes per query scan the nine
//                                runs together (lowest latency per query), the fit runs one query per thread on the
//                                neighbours staged in shared memory, tiles are assigned statically and the per-CTA sums
//                                added in CTA order.  Above: ONE THREAD per query with pruning -- a row or cell whose
//                                nearest face is farther than the current fifth-best distance is never read (typically 6
//                                of the 27 cells remain) -- chunks of 32 queries handed out dynamically to the warps, no
//                                CTA barrier inside an iteration, and the sums accumulated in 128-bit fixed point so that
//                                they do not depend on which warp took which chunk (gn_search / gn_chunk / GnFx).
//                                Measured on 1 M stack points against a 20 M-point map: 356 us per iteration against 496.
//
// Exactness: neighbours are ordered by (d2 ascending, original index ascending), d2 = ((dx*dx)+(dy*dy))+(dz*dz) in fp32
// without contraction; only neighbours with d2 < 1 m^2 can take part in an accepted correspondence (LM:762, 869), so the
// 27-cell search is exact for everything the reference uses.  The fit, the rows and the solve are the operations of the
// host code in the same order (lg_linalg.cuh, lg_libm.cuh); the sums are exact float products added in double in a fixed
// order and rounded once.
#include <cooperative_groups.h>

#include "lg_linalg.cuh"
#include "lg_map.h"
#include "lg_reduce.cuh"

namespace {

// ------------------------------------------------------------------------------------------------ CSR grid build
struct CsrJob {
  CsrGridD g[2];
  const float4* pts[2];
  int n[2];
  unsigned int* tab;  // both tables, back to back: [guard, cells of grid 0 ..., guard, cells of grid 1 ...]
  int off[2];         // first entry (the guard) of each grid's table inside `tab`
  float4* sorted;     // both grids' points in cell order, grid 0 first
};

__device__ __forceinline__ int csr_axis(float v, float origin, int n) {
  // cell coordinate clamped into the box: clamping is monotone and 1-Lipschitz, so two points less than 1 m apart along
  // an axis still land in the same or in adjacent cells -- a point outside the box is found from outside the box
  return (int)fminf(fmaxf(floorf(v) - origin, 0.f), (float)(n - 1));
}
__device__ __forceinline__ int csr_cell(const CsrGridD& g, float4 p) {
  const int cx = csr_axis(p.x, g.x0, g.nx), cy = csr_axis(p.y, g.y0, g.ny), cz = csr_axis(p.z, g.z0, g.nz);
  // entry 0 of a table is the guard in front, entry 1 + padded index belongs to the cell; a row is padded by one empty
  // cell on either side, so the real cell cx has padded index row + cx + 1
  return (cz * g.ny + cy) * g.nxp + cx + 2;
}

// min / max cell coordinates of a cloud (stage-level API only: the mapping node knows its cubes' box)
__global__ void csr_bbox_kernel(const float4* __restrict__ p0, int n0, const float4* __restrict__ p1, int n1, int* __restrict__ bb /* [2][6] */) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int w = i >= n0;
  const int k = w ? i - n0 : i;
  int lo[3] = {INT_MAX, INT_MAX, INT_MAX}, hi[3] = {INT_MIN, INT_MIN, INT_MIN};
  if (k < (w ? n1 : n0)) {
    const float4 p = w ? p1[k] : p0[k];
    const float c[3] = {p.x, p.y, p.z};
#pragma unroll
    for (int a = 0; a < 3; a++) {
      const int v = (int)fminf(fmaxf(floorf(c[a]), -1.0e6f), 1.0e6f);  // NaN -> lower bound, wild values clamped
      lo[a] = hi[a] = v;
    }
  }
  // a warp holds points of one cloud except at the seam; reduce per cloud with a match mask
  const unsigned int m = __match_any_sync(0xffffffffu, w);
#pragma unroll
  for (int a = 0; a < 3; a++) {
    lo[a] = __reduce_min_sync(m, lo[a]);
    hi[a] = __reduce_max_sync(m, hi[a]);
  }
  if ((threadIdx.x & 31) == __ffs(m) - 1 && lo[0] != INT_MAX) {
#pragma unroll
    for (int a = 0; a < 3; a++) {
      atomicMin(&bb[w * 6 + a], lo[a]);
      atomicMax(&bb[w * 6 + 3 + a], hi[a]);
    }
  }
}

__global__ void csr_count_kernel(CsrJob J, int split) {
  const int w = (int)blockIdx.x >= split;
  const int i = ((int)blockIdx.x - (w ? split : 0)) * blockDim.x + threadIdx.x;
  if (i >= J.n[w]) return;
  atomicAdd(&J.tab[J.off[w] + csr_cell(J.g[w], J.pts[w][i])], 1u);
}
__global__ void csr_fill_kernel(CsrJob J, int split) {
  const int w = (int)blockIdx.x >= split;
  const int i = ((int)blockIdx.x - (w ? split : 0)) * blockDim.x + threadIdx.x;
  if (i >= J.n[w]) return;
  const float4 p = J.pts[w][i];
  const unsigned int pos = atomicAdd(&J.tab[J.off[w] + csr_cell(J.g[w], p)], 1u);
  J.sorted[pos] = make_float4(p.x, p.y, p.z, __int_as_float(i));
}

// Exclusive scan of an unsigned array in place: tile sums, scan of the tile sums by one CTA, tile scan + offset.
constexpr int SCAN_NT = 1024, SCAN_ITEMS = 8, SCAN_TILE = SCAN_NT * SCAN_ITEMS;
__global__ void __launch_bounds__(SCAN_NT) scan_sums_kernel(const unsigned int* __restrict__ a, size_t len, unsigned int* __restrict__ sums) {
  __shared__ int s_w[SCAN_NT / 32 + 2];
  const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
  unsigned int s = 0;
  if (base + SCAN_ITEMS <= len) {
    const uint4 u = *reinterpret_cast<const uint4*>(a + base), v = *reinterpret_cast<const uint4*>(a + base + 4);
    s = u.x + u.y + u.z + u.w + v.x + v.y + v.z + v.w;
  } else {
    for (int k = 0; k < SCAN_ITEMS; k++)
      if (base + k < len) s += a[base + k];
  }
  int tot;
  block_excl_scan<SCAN_NT>((int)s, &tot, s_w);
  if (threadIdx.x == 0) sums[blockIdx.x] = (unsigned int)tot;
}
__global__ void __launch_bounds__(SCAN_NT) scan_top_kernel(unsigned int* __restrict__ sums, int nb) {
  __shared__ int s_w[SCAN_NT / 32 + 2];
  unsigned int carry = 0;
  for (int b0 = 0; b0 < nb; b0 += SCAN_NT) {
    const int b = b0 + threadIdx.x;
    const unsigned int v = b < nb ? sums[b] : 0u;
    int tot;
    const int ex = block_excl_scan<SCAN_NT>((int)v, &tot, s_w);
    if (b < nb) sums[b] = carry + (unsigned int)ex;
    carry += (unsigned int)tot;
  }
}
__global__ void __launch_bounds__(SCAN_NT) scan_apply_kernel(unsigned int* __restrict__ a, size_t len, const unsigned int* __restrict__ sums) {
  __shared__ int s_w[SCAN_NT / 32 + 2];
  const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
  unsigned int v[SCAN_ITEMS];
  if (base + SCAN_ITEMS <= len) {
    const uint4 u = *reinterpret_cast<const uint4*>(a + base), w = *reinterpret_cast<const uint4*>(a + base + 4);
    v[0] = u.x; v[1] = u.y; v[2] = u.z; v[3] = u.w; v[4] = w.x; v[5] = w.y; v[6] = w.z; v[7] = w.w;
  } else {
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) v[k] = base + k < len ? a[base + k] : 0u;
  }
  unsigned int s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) s += v[k];
  int tot;
  unsigned int run = sums[blockIdx.x] + (unsigned int)block_excl_scan<SCAN_NT>((int)s, &tot, s_w);
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    const unsigned int t = v[k];
    v[k] = run;
    run += t;
  }
  if (base + SCAN_ITEMS <= len) {
    *reinterpret_cast<uint4*>(a + base) = make_uint4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<uint4*>(a + base + 4) = make_uint4(v[4], v[5], v[6], v[7]);
  } else {
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++)
      if (base + k < len) a[base + k] = v[k];
  }
}

// ------------------------------------------------------------------------------------------------ the iteration
// LM:244-262 with the six sin/cos values of the pose (libm-exact: host sinf/cosf or their device ports).
__device__ __forceinline__ float4 gn_assoc_to_map(const float* t, const float* sc, float4 pi) {
  const float srx = sc[0], crx = sc[1], sry = sc[2], cry = sc[3], srz = sc[4], crz = sc[5];
  float x1 = crz * pi.x - srz * pi.y;
  float y1 = srz * pi.x + crz * pi.y;
  float z1 = pi.z;
  float x2 = x1;
  float y2 = crx * y1 - srx * z1;
  float z2 = srx * y1 + crx * z1;
  float4 po;
  po.x = cry * x2 + sry * z2 + t[3];
  po.y = y2 + t[4];
  po.z = -sry * x2 + cry * z2 + t[5];
  po.w = pi.w;
  return po;
}

__device__ __forceinline__ void gn_line_coeff(float x0, float y0, float z0, float x1, float y1, float z1, float x2, float y2, float z2, float& la,
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

// LM:763-964 for one stack point and its five neighbours: the row of A and b, or nothing.
__device__ __forceinline__ bool gn_fit_row(const float* sc, bool is_c, float4 ori, float4 sel, const float* px, const float* py, const float* pz,
                                           float* a, float* b) {
  float4 coef;
  bool keep = false;
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
    lg_jacobi_eigen3(A1, D1, V1);
    if (D1[0] > 3 * D1[1]) {
      float x1 = (float)(cx + 0.1 * V1[0]), y1 = (float)(cy + 0.1 * V1[1]), z1 = (float)(cz + 0.1 * V1[2]);
      float x2 = (float)(cx - 0.1 * V1[0]), y2 = (float)(cy - 0.1 * V1[1]), z2 = (float)(cz - 0.1 * V1[2]);
      float la, lb, lc, ld2;
      gn_line_coeff(sel.x, sel.y, sel.z, x1, y1, z1, x2, y2, z2, la, lb, lc, ld2);
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
  if (!keep) return false;
  // LM:940-964
  const float srx = sc[0], crx = sc[1], sry = sc[2], cry = sc[3], srz = sc[4], crz = sc[5];
  const float4 p = ori, c = coef;
  a[0] = (crx * sry * srz * p.x + crx * crz * sry * p.y - srx * sry * p.z) * c.x + (-srx * srz * p.x - crz * srx * p.y - crx * p.z) * c.y +
         (crx * cry * srz * p.x + crx * cry * crz * p.y - cry * srx * p.z) * c.z;
  a[1] = ((cry * srx * srz - crz * sry) * p.x + (sry * srz + cry * crz * srx) * p.y + crx * cry * p.z) * c.x +
         ((-cry * crz - srx * sry * srz) * p.x + (cry * srz - crz * srx * sry) * p.y - crx * sry * p.z) * c.z;
  a[2] = ((crz * srx * sry - cry * srz) * p.x + (-cry * crz - srx * sry * srz) * p.y) * c.x + (crx * crz * p.x - crx * srz * p.y) * c.y +
         ((sry * srz + cry * crz * srx) * p.x + (crz * sry - cry * srx * srz) * p.y) * c.z;
  a[3] = c.x;
  a[4] = c.y;
  a[5] = c.z;
  *b = -c.w;
  return true;
}

// LOAM_GN_DEBUG=1: nanosecond stamps of the phases of every iteration (dbg[iteration][8]), printed by the host after the
// launch -- how the phase costs quoted in DESIGN.md were measured
#define GN_STAMP(k, cond)                                                              \
  do {                                                                                 \
    if (A.dbg != nullptr && (cond)) {                                                  \
      unsigned long long t_;                                                           \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                            \
      A.dbg[(size_t)(iter - A.it0) * 8 + (k)] = t_;                                    \
    }                                                                                  \
  } while (0)
constexpr int GN_NT = 256;           // threads per CTA
constexpr int GN_TILE = 256;         // largest number of stack points per CTA step (one per thread in the fit phase)
constexpr unsigned long long GN_EMPTY = ~0ull;

// Sorted insertion of (key, payload) into an ascending 5-list held in registers, without branches on the position.
#define GN_INSERT(c, pc)                                                                         \
  do {                                                                                           \
    const bool l3 = (c) < k3, l2 = (c) < k2, l1 = (c) < k1, l0 = (c) < k0;                        \
    k4 = l3 ? k3 : (c); p4 = l3 ? p3 : (pc);                                                      \
    k3 = l2 ? k2 : (l3 ? (c) : k3); p3 = l2 ? p2 : (l3 ? (pc) : p3);                              \
    k2 = l1 ? k1 : (l2 ? (c) : k2); p2 = l1 ? p1 : (l2 ? (pc) : p2);                              \
    k1 = l0 ? k0 : (l1 ? (c) : k1); p1 = l0 ? p0 : (l1 ? (pc) : p1);                              \
    k0 = l0 ? (c) : k0; p0 = l0 ? (pc) : p0;                                                      \
  } while (0)

struct GnCommon {
  int wcount[GN_NT / 32];
  double acc[GN_NT / 32][28];  // per-warp sums of the rows of this CTA
  double tot[28];
  float T[6], sc[6];
  int done, stop;
  bool last;
};
struct GnSharedL : GnCommon {  // SUB lanes per query
  float4 ori[GN_TILE];        // stack point (sensor frame)
  float4 sel[GN_TILE];        // the same in the map frame under the current pose (LM:756, 865)
  float nb[GN_TILE][5][3];    // its five neighbours
  float row[GN_TILE][8];      // its row of A (6), b, and 1 / 0 = kept (LM:940-967)
  unsigned short list[GN_TILE];  // queries of the tile this rank evaluates (owner rule), in order
  unsigned char ok[GN_TILE];  // five neighbours within 1 m found
  int n_own;
};
struct GnSharedT : GnCommon {  // one thread per query
  float row[GN_TILE][8];
};

// rows (y, z offsets) of the 3 x 3 x 3 neighbourhood in the order they are visited: own row, the four that share a face
// with it, the four that share an edge -- nearer rows first, so the fifth-best distance shrinks early
__constant__ signed char GN_ROW_OY[9] = {0, -1, 1, 0, 0, -1, 1, -1, 1};
__constant__ signed char GN_ROW_OZ[9] = {0, 0, 0, -1, 1, -1, -1, 1, 1};

// The warp's rows -> its 28 sums: the rows go through shared memory; lane L < 28 then adds "its" product over the 32 rows
// in row order (exact float products, double sums): 21 upper-triangle terms of AtA, 6 of AtB, the row count.
__device__ __forceinline__ double gn_warp_rows(float (*rows)[8], int lane, bool keep, const float* a, float b) {
  if (!__any_sync(0xffffffffu, keep)) return 0.0;
  double sum = 0.0;
  float* mine = rows[lane];
#pragma unroll
  for (int i = 0; i < 6; i++) mine[i] = keep ? a[i] : 0.f;
  mine[6] = keep ? b : 0.f;
  mine[7] = keep ? 1.f : 0.f;
  __syncwarp();
  if (lane < 28) {
    // lane -> (i, j): 0..20 upper triangle row-major, 21..26 (i, 6) = AtB, 27 (7, 7) = count
    int i = 0, j = lane;
    if (lane < 21) {
      int rem = lane;
#pragma unroll
      for (int r = 0; r < 6; r++)
        if (rem >= 6 - r && i == r) {
          rem -= 6 - r;
          i = r + 1;
        }
      j = i + rem;
    } else if (lane < 27) {
      i = lane - 21;
      j = 6;
    } else {
      i = 7;
      j = 7;
    }
#pragma unroll 8
    for (int r = 0; r < 32; r++) sum += (double)rows[r][i] * (double)rows[r][j];
  }
  __syncwarp();
  return sum;
}

// One thread, one stack point: exact 5-NN over the 27 cells around it with pruning, fit, row.  A cell (or a whole row of
// three) is skipped when no point in it can enter the list: the bound below is the distance to the cell's nearest face,
// formed with the SAME rounded operations in the same order as a point's distance, and fp32 rounding is monotone -- so
// every point of the cell has a computed d2 >= the bound, and skipping on `bound >= 1` (LM:762, 869 accept d2 < 1 only) or
// `bound > fifth-best d2` never drops a point the exhaustive search would have kept (an equal d2 is NOT skipped: the
// smaller original index wins ties).
struct GnNbr {
  unsigned long long k0, k1, k2, k3, k4;
  unsigned int p0, p1, p2, p3, p4;
};
struct GnGrid {  // one cloud's grid, warp-uniform
  const float4* pts;
  const uint4* E4;  // the run-end table read four entries at a time (16-byte aligned)
  float x0, y0, z0;
  int nx, ny, nz, nxp;
};
// two points with one 256-bit load (sm_100: LDG.E.256); p must be 32-byte aligned
__device__ __forceinline__ void gn_ld2(const float4* p, float4& a, float4& b) {
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
               : "l"(p));
}
__device__ __forceinline__ void gn_search(const GnGrid& g, float4 sel, GnNbr& N) {
  unsigned long long k0 = GN_EMPTY, k1 = GN_EMPTY, k2 = GN_EMPTY, k3 = GN_EMPTY, k4 = GN_EMPTY;
  unsigned int p0 = 0, p1 = 0, p2 = 0, p3 = 0, p4 = 0;
  const float4* __restrict__ pts = g.pts;
  const int cx = csr_axis(sel.x, g.x0, g.nx), cy = csr_axis(sel.y, g.y0, g.ny), cz = csr_axis(sel.z, g.z0, g.nz);
  // distances to the six faces of the own cell (the faces between cells are exact integers)
  const float fxl = sel.x - (g.x0 + (float)cx), fxh = (g.x0 + (float)(cx + 1)) - sel.x;
  const float fyl = sel.y - (g.y0 + (float)cy), fyh = (g.y0 + (float)(cy + 1)) - sel.y;
  const float fzl = sel.z - (g.z0 + (float)cz), fzh = (g.z0 + (float)(cz + 1)) - sel.z;
  const float fxl2 = fxl * fxl, fxh2 = fxh * fxh;
  auto prune = [&](float lb) { return lb >= 1.0f || __float_as_uint(lb) > (unsigned int)(k4 >> 32); };
  // a row's four table entries E[idx .. idx + 3] sit in two aligned 16-byte words; they are requested one row AHEAD of the
  // scan that needs them, so the look-up of the next row travels while this row's points are compared
  uint4 nA = make_uint4(0u, 0u, 0u, 0u), nB = nA;
  unsigned int n_sh = 0;
  float n_ey2 = 0.f, n_ez2 = 0.f;
  bool n_ok = false;
  auto request = [&](int r) {
    const int oy = GN_ROW_OY[r], oz = GN_ROW_OZ[r];
    const int ry = cy + oy, rz = cz + oz;
    const float ey = oy == 0 ? 0.f : (oy < 0 ? fyl : fyh), ez = oz == 0 ? 0.f : (oz < 0 ? fzl : fzh);
    n_ey2 = ey * ey;
    n_ez2 = ez * ez;
    n_ok = ry >= 0 && ry < g.ny && rz >= 0 && rz < g.nz && !prune(n_ey2 + n_ez2);
    if (n_ok) {
      const size_t idx = (size_t)(rz * g.ny + ry) * g.nxp + cx;
      const uint4* a = g.E4 + (idx >> 2);
      n_sh = (unsigned int)idx & 3u;
      nA = __ldg(a);
      nB = __ldg(a + 1);
    }
  };
  request(0);
#pragma unroll 1
  for (int r = 0; r < 9; r++) {
    const bool ok = n_ok;
    const uint4 A0 = nA, A1 = nB;
    const unsigned int sh = n_sh;
    const float ey2 = n_ey2, ez2 = n_ez2;
    if (r + 1 < 9) request(r + 1);
    if (!ok || prune(ey2 + ez2)) continue;
    const unsigned int e0 = sh == 0 ? A0.x : (sh == 1 ? A0.y : (sh == 2 ? A0.z : A0.w));
    const unsigned int e1 = sh == 0 ? A0.y : (sh == 1 ? A0.z : (sh == 2 ? A0.w : A1.x));
    const unsigned int e2 = sh == 0 ? A0.z : (sh == 1 ? A0.w : (sh == 2 ? A1.x : A1.y));
    const unsigned int e3 = sh == 0 ? A0.w : (sh == 1 ? A1.x : (sh == 2 ? A1.y : A1.z));
    // own row: the query's own cell first, then -- against the tightened bound -- the cells left and right of it; the other
    // rows: their (up to) three cells are one contiguous run, trimmed by the bound of the moment
    const int nseg = r == 0 ? 3 : 1;
#pragma unroll 1
    for (int ci = 0; ci < nseg; ci++) {
      unsigned int s, e;
      if (r == 0) {
        s = ci == 0 ? e1 : (ci == 1 ? e0 : e2);
        e = ci == 0 ? e2 : (ci == 1 ? e1 : e3);
        if (ci > 0 && prune(((ci == 1 ? fxl2 : fxh2) + ey2) + ez2)) continue;
      } else {
        s = (e0 < e1 && !prune((fxl2 + ey2) + ez2)) ? e0 : e1;
        e = (e2 < e3 && !prune((fxh2 + ey2) + ez2)) ? e3 : e2;
      }
      for (unsigned int p = s & ~1u; p < e; p += 4) {
        float4 c[4];
        gn_ld2(pts + p, c[0], c[1]);
        if (p + 2 < e) gn_ld2(pts + p + 2, c[2], c[3]);
#pragma unroll
        for (int u = 0; u < 4; u++)
          if (p + u >= s && p + u < e) {
            const float d2 = lg_sqdist(c[u].x, c[u].y, c[u].z, sel.x, sel.y, sel.z);
            if (d2 < 1.0f) {
              const unsigned long long key = lg_pack_nbr(d2, __float_as_int(c[u].w));
              if (key < k4) GN_INSERT(key, p + u);
            }
          }
      }
    }
  }
  N.k0 = k0; N.k1 = k1; N.k2 = k2; N.k3 = k3; N.k4 = k4;
  N.p0 = p0; N.p1 = p1; N.p2 = p2; N.p3 = p3; N.p4 = p4;
}

// Order-independent accumulation of the chunk sums: every chunk's 28 sums (32 rows added in row order: deterministic) are
// converted to 128-bit fixed point (64 fractional bits, truncation) and added as four 32-bit limbs into 64-bit integers --
// integer addition is associative, so the totals do not depend on which warp took which chunk or on the order the warps
// finish in, although the chunks are handed out dynamically.
struct GnFx {
  unsigned long long l0, l1, l2, l3;
  __device__ __forceinline__ void clear() { l0 = l1 = l2 = l3 = 0ull; }
  __device__ __forceinline__ void add(double s) {
    const long long hi = __double2ll_rd(s);
    const double frac = s - (double)hi;  // in [0, 1), exact
    const unsigned long long lo = __double2ull_rd(frac * 18446744073709551616.0);
    l0 += lo & 0xffffffffull;
    l1 += lo >> 32;
    l2 += (unsigned long long)hi & 0xffffffffull;
    l3 += (unsigned long long)(hi >> 32);  // arithmetic shift: the sign lives in the top limb
  }
};
__device__ __forceinline__ double gn_fx_total(unsigned long long l0, unsigned long long l1, unsigned long long l2, unsigned long long l3) {
  const unsigned __int128 v = (unsigned __int128)l0 + ((unsigned __int128)l1 << 32) + ((unsigned __int128)l2 << 64) + ((unsigned __int128)l3 << 96);
  const long long hi = (long long)(unsigned long long)(v >> 64);
  const unsigned long long lo = (unsigned long long)v;
  return (double)hi + (double)lo * 5.421010862427522e-20;  // 2^-64
}

// One chunk = 32 consecutive points of ONE stack cloud, one per lane: transform, owner rule, search, fit, row; the lanes
// < 28 return with their sum over the chunk's rows added to fx.
__device__ __forceinline__ void gn_chunk(const MapGnArgs& A, const float* T, const float* sc, int chunk, int n_chunks_c, int lane, float (*rows)[8],
                                         GnFx& fx) {
  const bool is_c = chunk < n_chunks_c;  // warp-uniform
  const int q = (is_c ? chunk : chunk - n_chunks_c) * 32 + lane;
  const int n_cloud = is_c ? A.n_cs : A.n_ss;
  const float4* __restrict__ stack = is_c ? A.cstack : A.sstack;
  GnGrid g;  // field by field: a reference picked at run time would force both parameter structs into local memory
  g.pts = is_c ? A.gc.sorted : A.gs.sorted;
  g.E4 = reinterpret_cast<const uint4*>(is_c ? A.gc.E : A.gs.E);
  g.x0 = is_c ? A.gc.x0 : A.gs.x0; g.y0 = is_c ? A.gc.y0 : A.gs.y0; g.z0 = is_c ? A.gc.z0 : A.gs.z0;
  g.nx = is_c ? A.gc.nx : A.gs.nx; g.ny = is_c ? A.gc.ny : A.gs.ny; g.nz = is_c ? A.gc.nz : A.gs.nz;
  g.nxp = is_c ? A.gc.nxp : A.gs.nxp;
  float a[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, b = 0.f;
  bool keep = false;
  if (q < n_cloud) {
    const float4 ori = __ldg(&stack[q]);
    const float4 sel = gn_assoc_to_map(T, sc, ori);
    if (sel.x >= A.slab_lo && sel.x < A.slab_hi) {  // owner rule of a sharded map
      GnNbr N;
      gn_search(g, sel, N);
      const bool found = N.k4 != GN_EMPTY;
      if (A.nbr != nullptr) {
        int* o = A.nbr + (size_t)(is_c ? q : A.n_cs + q) * 5;
        o[0] = found ? (int)(unsigned int)N.k0 : -1;
        o[1] = found ? (int)(unsigned int)N.k1 : -1;
        o[2] = found ? (int)(unsigned int)N.k2 : -1;
        o[3] = found ? (int)(unsigned int)N.k3 : -1;
        o[4] = found ? (int)(unsigned int)N.k4 : -1;
      }
      if (found) {
        const float4* __restrict__ pts = g.pts;
        const float4 n0 = __ldg(&pts[N.p0]), n1 = __ldg(&pts[N.p1]), n2 = __ldg(&pts[N.p2]), n3 = __ldg(&pts[N.p3]), n4 = __ldg(&pts[N.p4]);
        const float px[5] = {n0.x, n1.x, n2.x, n3.x, n4.x}, py[5] = {n0.y, n1.y, n2.y, n3.y, n4.y}, pz[5] = {n0.z, n1.z, n2.z, n3.z, n4.z};
        keep = gn_fit_row(sc, is_c, ori, sel, px, py, pz, a, &b);
      }
    }
  }
  const double rs = gn_warp_rows(rows, lane, keep, a, b);
  if (lane < 28 && rs != 0.0) fx.add(rs);
}

// SUB = lanes per query in the search phase.  8 (or 4): the lanes of a group scan the nine runs together -- lowest latency
// per query, for stacks too small to fill the GPU (one sweep against the local map).  1: one thread per query with cell
// pruning (gn_search) -- an order of magnitude fewer instructions per query, for large stacks.
template <int SUB>
struct GnSharedOf { typedef GnSharedL type; };
template <>
struct GnSharedOf<1> { typedef GnSharedT type; };

template <int SUB, int MINB>
__global__ void __launch_bounds__(GN_NT, MINB) map_gn_kernel(MapGnArgs A) {
  __shared__ typename GnSharedOf<SUB>::type S;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int sub = lane & (SUB - 1), grp = lane / SUB;
  const int nq = A.n_cs + A.n_ss;
  const int tq = A.tile;  // stack points per CTA step: 256 when there is work for every SM, down to 32 (one search step) otherwise
  const int ntiles = (nq + tq - 1) / tq;
  if (tid < 6) {
    S.T[tid] = A.T[tid];
    S.sc[tid] = A.sc[tid];
  }
  unsigned int gen = 0;  // grid-barrier generations this launch has passed
  // generation counter at launch: read before this CTA's first ticket, i.e. before the first iteration can complete
  const unsigned int gen0 = tid == 0 ? *((volatile unsigned int*)A.gen) : 0u;
  __syncthreads();
  for (int iter = A.it0; iter < A.it1; iter++) {
    float T[6], sc[6];
    GN_STAMP(0, blockIdx.x == 0 && tid == 0);
#pragma unroll
    for (int i = 0; i < 6; i++) {
      T[i] = S.T[i];
      sc[i] = S.sc[i];
    }
    if (SUB != 1 && lane < 28) S.acc[w][lane] = 0.0;
    if constexpr (SUB == 1) {
      // chunks of 32 stack points are handed out dynamically, one per warp at a time (the next ticket is drawn before the
      // current chunk is worked on, so its round trip to L2 is hidden); no CTA barrier inside the iteration
      const int n_chunks_c = (A.n_cs + 31) >> 5, n_chunks = n_chunks_c + ((A.n_ss + 31) >> 5);
      GnFx fx;
      fx.clear();
      int next = 0;
      if (lane == 0) next = (int)atomicAdd(A.work, 1u);
      for (;;) {
        const int chunk = __shfl_sync(0xffffffffu, next, 0);
        if (chunk >= n_chunks) break;
        if (lane == 0) next = (int)atomicAdd(A.work, 1u);
        gn_chunk(A, T, sc, chunk, n_chunks_c, lane, &S.row[w * 32], fx);
      }
      if (lane < 28 && (fx.l0 | fx.l1 | fx.l2 | fx.l3) != 0ull) {
        unsigned long long* G = A.fx + lane * 4;
        atomicAdd(G + 0, fx.l0);
        atomicAdd(G + 1, fx.l1);
        atomicAdd(G + 2, fx.l2);
        atomicAdd(G + 3, fx.l3);
      }
    } else {
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      // ---- phase 0: transform, owner rule, ordered compaction of the owned queries
      const int q = tile * tq + tid;
      bool own = false;
      if (tid < tq && q < nq) {
        const float4 ori = q < A.n_cs ? A.cstack[q] : A.sstack[q - A.n_cs];
        const float4 sel = gn_assoc_to_map(T, sc, ori);
        S.ori[tid] = ori;
        S.sel[tid] = sel;
        own = sel.x >= A.slab_lo && sel.x < A.slab_hi;
      }
      S.ok[tid] = 0;
      const unsigned int bal = __ballot_sync(0xffffffffu, own);
      if (lane == 0) S.wcount[w] = __popc(bal);
      __syncthreads();
      {
        int base = 0, total = 0;
#pragma unroll
        for (int k = 0; k < GN_NT / 32; k++) {
          const int c = S.wcount[k];
          base += k < w ? c : 0;
          total += c;
        }
        if (own) S.list[base + __popc(bal & ((1u << lane) - 1u))] = (unsigned short)tid;
        if (tid == 0) S.n_own = total;
      }
      __syncthreads();
      // ---- phase 1: exact 5-NN, SUB lanes per query (32 / SUB queries per warp and step)
      const int n_own = S.n_own;
      constexpr int QPW = 32 / SUB;            // queries per warp and step
      constexpr int RPL = (9 + SUB - 1) / SUB;  // rows a lane looks up
      const int gbase = grp * SUB;
      // row look-ups of a step: lane `sub` takes rows sub, sub + SUB, ... of its query; they are issued one step AHEAD of
      // the scan that needs them, so their latency hides behind the previous step's point loads
      unsigned int lo[RPL], hi[RPL];
      float4 sel_n = make_float4(0.f, 0.f, 0.f, 0.f);
      int lq_n = 0;
      bool active_n = false;
      const float4* pts_n = nullptr;
      auto lookup = [&](int e0) {
        const int e = e0 + grp;
        active_n = e < n_own;
        lq_n = active_n ? (int)S.list[e] : 0;
        sel_n = S.sel[lq_n];
        const CsrGridD& g = (tile * tq + lq_n < A.n_cs) ? A.gc : A.gs;
        pts_n = g.sorted;
        const int cx = csr_axis(sel_n.x, g.x0, g.nx), cy = csr_axis(sel_n.y, g.y0, g.ny), cz = csr_axis(sel_n.z, g.z0, g.nz);
#pragma unroll
        for (int t = 0; t < RPL; t++) {
          const int row = sub + t * SUB;
          const int ry = cy + (row % 3) - 1, rz = cz + (row / 3) - 1;
          lo[t] = hi[t] = 0u;
          if (active_n && row < 9 && ry >= 0 && ry < g.ny && rz >= 0 && rz < g.nz) {
            const unsigned int* r = g.E + (size_t)(rz * g.ny + ry) * g.nxp + cx;
            lo[t] = __ldg(r);
            hi[t] = __ldg(r + 3);
          }
        }
      };
      if (w * QPW < n_own) lookup(w * QPW);
      for (int e0 = w * QPW; e0 < n_own; e0 += QPW * (GN_NT / 32)) {  // warp-uniform trip count: the groups of a warp stay together
        const bool active = active_n;
        const int lq = lq_n;
        const float4 sel = sel_n;
        const float4* __restrict__ pts = pts_n;
        unsigned int rl[9], rh[9];
        int rounds = 0;
#pragma unroll
        for (int j = 0; j < 9; j++) {
          const unsigned int a0 = __shfl_sync(0xffffffffu, lo[j / SUB], gbase + (j % SUB));
          rh[j] = __shfl_sync(0xffffffffu, hi[j / SUB], gbase + (j % SUB));
          rounds = max(rounds, (int)(rh[j] - a0 + SUB - 1) / SUB);
          rl[j] = a0 + sub;
        }
        if (e0 + QPW * (GN_NT / 32) < n_own) lookup(e0 + QPW * (GN_NT / 32));
        unsigned long long k0 = GN_EMPTY, k1 = GN_EMPTY, k2 = GN_EMPTY, k3 = GN_EMPTY, k4 = GN_EMPTY;
        unsigned int p0 = 0, p1 = 0, p2 = 0, p3 = 0, p4 = 0;
        // every lane takes points sub, sub + SUB, ... of each of the nine runs; the (up to) nine loads of a round are issued
        // together before the first distance is formed, so a lane has nine L2 / DRAM requests in flight instead of one
        for (int rd = 0; rd < rounds; rd++) {
          float4 pf[9];
#pragma unroll
          for (int j = 0; j < 9; j++)
            if (rl[j] < rh[j]) pf[j] = __ldg(&pts[rl[j]]);
#pragma unroll
          for (int j = 0; j < 9; j++) {
            if (rl[j] < rh[j]) {
              const float d2 = lg_sqdist(pf[j].x, pf[j].y, pf[j].z, sel.x, sel.y, sel.z);
              if (d2 < 1.0f) {
                const unsigned long long c = lg_pack_nbr(d2, __float_as_int(pf[j].w));
                if (c < k4) GN_INSERT(c, rl[j]);
              }
            }
            rl[j] += SUB;
          }
        }
        // merge inside the group: five rounds of "smallest head of the lanes' sorted lists".  The minimum of the distance
        // bits is found with xor-shuffles (they stay inside the group for every group of the warp at once); an exact tie
        // between two lanes' heads (rare) is broken by the smaller original index the same way.
        unsigned int pos_a = 0, pos_b = 0;  // lane `sub` ends with the positions of neighbours sub and sub + SUB
        int idx_a = -1, idx_b = -1;
        bool found = false;
#pragma unroll
        for (int r = 0; r < 5; r++) {
          const unsigned int hi32 = (unsigned int)(k0 >> 32), lo32 = (unsigned int)k0;
          unsigned int m = hi32;
#pragma unroll
          for (int o = 1; o < SUB; o <<= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, o));
          bool cand = hi32 == m && m != 0xffffffffu;
          unsigned int bal = (__ballot_sync(0xffffffffu, cand) >> gbase) & ((1u << SUB) - 1u);
          if (__any_sync(0xffffffffu, (bal & (bal - 1u)) != 0u)) {  // some group of the warp has a tie on d2
            unsigned int ml = cand ? lo32 : 0xffffffffu;
#pragma unroll
            for (int o = 1; o < SUB; o <<= 1) ml = min(ml, __shfl_xor_sync(0xffffffffu, ml, o));
            cand = cand && lo32 == ml;
            bal = (__ballot_sync(0xffffffffu, cand) >> gbase) & ((1u << SUB) - 1u);
          }
          const int wl = gbase + (bal ? __ffs(bal) - 1 : 0);
          const unsigned int pos = __shfl_sync(0xffffffffu, p0, wl);
          const unsigned int idx = __shfl_sync(0xffffffffu, lo32, wl);
          if (sub == (r % SUB)) {
            if (r < SUB) {
              pos_a = pos;
              idx_a = (int)idx;
            } else {
              pos_b = pos;
              idx_b = (int)idx;
            }
          }
          if (r == 4) found = active && bal != 0u;
          if (cand) {
            k0 = k1; k1 = k2; k2 = k3; k3 = k4; k4 = GN_EMPTY;
            p0 = p1; p1 = p2; p2 = p3; p3 = p4;
          }
        }
        if (sub < 5 && found) {
          const float4 p = __ldg(&pts[pos_a]);
          S.nb[lq][sub][0] = p.x;
          S.nb[lq][sub][1] = p.y;
          S.nb[lq][sub][2] = p.z;
        }
        if (SUB < 5 && sub + SUB < 5 && found) {
          const float4 p = __ldg(&pts[pos_b]);
          S.nb[lq][sub + SUB][0] = p.x;
          S.nb[lq][sub + SUB][1] = p.y;
          S.nb[lq][sub + SUB][2] = p.z;
        }
        if (A.nbr != nullptr && active) {
          int* o = A.nbr + (size_t)(tile * tq + lq) * 5;
          if (sub < 5) o[sub] = found ? idx_a : -1;
          if (SUB < 5 && sub + SUB < 5) o[sub + SUB] = found ? idx_b : -1;
        }
        if (sub == 0 && found) S.ok[lq] = 1;
      }
      __syncthreads();
      if (tile == (int)blockIdx.x) GN_STAMP(6, blockIdx.x == 0 && tid == 0);
      // ---- phase 2: fit + Jacobian row, one query per thread; the warp's 28 sums go to its shared accumulator
      if (w * 32 < tq) {  // warp-uniform: only the warps that hold queries of this tile
        float a[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, b = 0.f;
        bool keep = false;
        if (S.ok[tid]) {
          float px[5], py[5], pz[5];
#pragma unroll
          for (int j = 0; j < 5; j++) {
            px[j] = S.nb[tid][j][0];
            py[j] = S.nb[tid][j][1];
            pz[j] = S.nb[tid][j][2];
          }
          keep = gn_fit_row(sc, q < A.n_cs, S.ori[tid], S.sel[tid], px, py, pz, a, &b);
        }
        const double rs = gn_warp_rows(&S.row[w * 32], lane, keep, a, b);
        if (lane < 28) S.acc[w][lane] += rs;
      }
      __syncthreads();
    }
    }
    GN_STAMP(1, blockIdx.x == 0 && tid == 0);
    // ---- grid reduction.  Lanes-per-query layouts: CTA partial -> global, the last CTA adds them in CTA order; one thread
    // per query: the fixed-point totals are already in global memory, the last CTA converts them
    __syncthreads();
    if constexpr (SUB != 1) {
      if (tid < 28) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < GN_NT / 32; k++) s += S.acc[k][tid];
        A.partials[(size_t)blockIdx.x * 28 + tid] = s;
      }
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      const unsigned int t = atomicAdd(A.ticket, 1u);
      S.last = (t == gridDim.x - 1);
      if (S.last) GN_STAMP(2, true);
    }
    __syncthreads();
    if (S.last) {
      __threadfence();
      double s = 0.0;
      if constexpr (SUB == 1) {
        if (tid < 28) {
          unsigned long long* G = A.fx + tid * 4;
          s = gn_fx_total(__ldcg(G + 0), __ldcg(G + 1), __ldcg(G + 2), __ldcg(G + 3));
        }
        __syncthreads();
        if (tid < 28 * 4) A.fx[tid] = 0ull;  // every CTA has added its share (ticket): clean for the next iteration / launch
        if (tid == 0) *A.work = 0u;
      } else {
        if (lane < 28) {
          constexpr unsigned int W = GN_NT / 32;
          for (unsigned int b = w; b < gridDim.x; b += 4 * W) {
            double v[4];
#pragma unroll
            for (unsigned int u = 0; u < 4; u++) v[u] = b + u * W < gridDim.x ? __ldcg(&A.partials[(size_t)(b + u * W) * 28 + lane]) : 0.0;
#pragma unroll
            for (unsigned int u = 0; u < 4; u++) s += v[u];
          }
          S.acc[w][lane] = s;
        }
        __syncthreads();
        if (tid < 28) {
          s = 0.0;
#pragma unroll
          for (int k = 0; k < GN_NT / 32; k++) s += S.acc[k][tid];
        }
      }
      if (tid == 0) *A.ticket = 0u;
      if (A.px.world > 1) {
        // One-shot all-reduce over NVLink peer memory (see lg_reduce.cuh): store this rank's sums into every peer's
        // exchange buffer, raise the flag there, wait for the peers' flags here, add the partials in rank order.
        const int W = A.px.world, me = A.px.rank;
        const unsigned long long xs = A.px.xseq + (unsigned long long)(iter - A.it0);
        const size_t slot = ((size_t)(xs & 1ull) * LG_MAX_PEERS + me) * 32;
        if (tid < 28)
          for (int r = 0; r < W; r++) A.px.buf[r][slot + tid] = s;
        __threadfence_system();
        __syncthreads();
        if (tid < W) *((volatile unsigned long long*)(A.px.buf[tid] + LG_XCHG_FLAG_OFFSET) + me) = xs;
        if (tid < W) {
          volatile unsigned long long* f = (volatile unsigned long long*)(A.px.buf[me] + LG_XCHG_FLAG_OFFSET) + tid;
          const long long t0 = clock64();
          while (*f < xs) {
            if (clock64() - t0 > (1ll << 32)) {  // ~2 s: a peer died or never called; the host turns this into an error
              *A.px.timeout = 1;
              break;
            }
          }
        }
        __threadfence_system();
        __syncthreads();
        if (tid < 28) {
          s = 0.0;
          const volatile double* mine = A.px.buf[me] + (size_t)(xs & 1ull) * LG_MAX_PEERS * 32;
          for (int r = 0; r < W; r++) s += mine[(size_t)r * 32 + tid];
        }
      }
      if (tid < 28) S.tot[tid] = s;
      GN_STAMP(3, tid == 0);
      __syncthreads();
      if (!A.solve) {
        if (tid < 28) A.out[tid] = S.tot[tid];
      } else if (w == 0) {
        // ---- LM:929-932, 968-1017 on the device, same operations in the same order as the host code
        float X[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        float AtA[36], AtB[6];
        int n_sel;
        lg_unpack28(S.tot, AtA, AtB, &n_sel);
        const int solved = n_sel >= 50;
        float Tn[6];
#pragma unroll
        for (int i = 0; i < 6; i++) Tn[i] = T[i];
        if (solved) {
          lg_qr_solve6_warp(AtA, AtB, lane, X);
          // iteration 0: the eigen-decomposition / degeneracy test (LM:970-997) is checked by the host afterwards on the
          // sums published here; the device goes on as if it said "not degenerate" (the host re-runs the rare other case)
          int degenerate = A.degenerate;
          if (iter == 0) {
            degenerate = 0;
            if (lane < 28) A.out[lane] = S.tot[lane];
            if (lane == 0) A.out[40] = 1.0;
          }
          if (degenerate) {
            float X2[6];
            for (int i = 0; i < 6; i++) X2[i] = X[i];
            lg_gemm_dacc(A.matP, X2, X, 6, 6, 1);
          }
#pragma unroll
          for (int i = 0; i < 6; i++) Tn[i] = T[i] + X[i];
        }
        int small = 0;
        if (lane < 6) {
          const int k = lane >> 1;
          const float ang = k == 0 ? Tn[0] : (k == 1 ? Tn[1] : Tn[2]);
          S.sc[lane] = (lane & 1) ? lgm_cosf(ang) : lgm_sinf(ang);
          float tv = Tn[0];
#pragma unroll
          for (int i = 1; i < 6; i++) tv = lane == i ? Tn[i] : tv;
          S.T[lane] = tv;
        } else if (lane == 6) {
          const double r0 = X[0] * 180.0 / M_PI, r1 = X[1] * 180.0 / M_PI, r2 = X[2] * 180.0 / M_PI;
          small = (float)sqrt(r0 * r0 + r1 * r1 + r2 * r2) < 0.05;
        } else if (lane == 7) {
          const double t0 = X[3] * 100, t1 = X[4] * 100, t2 = X[5] * 100;
          small = (float)sqrt(t0 * t0 + t1 * t1 + t2 * t2) < 0.05;
        }
        const unsigned int both = __ballot_sync(0xffffffffu, small) & 0xc0u;
        __syncwarp();
        if (lane == 0) S.done = (solved && both == 0xc0u) ? 1 : 0;
        __syncwarp();
        // publish pose, sin / cos and the verdict for every CTA (and, at the end, for the host)
        if (lane < 6) {
          A.state[lane] = S.T[lane];
          A.state[6 + lane] = S.sc[lane];
        }
        if (lane == 0) {
          A.state[12] = __int_as_float(S.done);
          A.state[13] = __int_as_float(iter);
        }
      }
      __syncthreads();
      const bool stop = !A.solve || S.done || iter + 1 >= A.it1;
      if (stop) {  // host mailbox: results first, then the sequence word
        if (A.solve && tid < 6) A.out[32 + tid] = (double)S.T[tid];
        if (A.solve && tid == 6) A.out[38] = (double)iter;
        if (A.solve && tid == 7) A.out[39] = (double)S.done;
        if (A.seq != 0ull) {
          __threadfence_system();
          __syncthreads();
          if (tid == 0) {
            *((volatile unsigned long long*)(A.out + 31)) = A.seq;
            __threadfence_system();
          }
        }
      }
      __threadfence();
      __syncthreads();
      GN_STAMP(4, tid == 0);
      if (tid == 0) atomicAdd(A.gen, 1u);  // release the grid barrier
    }
    if (!A.solve) return;
    // ---- grid barrier: wait for the last CTA's verdict, pick up the new pose
    gen++;
    if (tid == 0) {
      while (*((volatile unsigned int*)A.gen) - gen0 < gen) {
      }
      __threadfence();
    }
    __syncthreads();
    if (tid < 6) {
      S.T[tid] = __ldcg(&A.state[tid]);
      S.sc[tid] = __ldcg(&A.state[6 + tid]);
    }
    if (tid == 0) S.done = __float_as_int(__ldcg(&A.state[12]));
    __syncthreads();
    GN_STAMP(5, blockIdx.x == 0 && tid == 0);
    if (S.done) break;
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------------------ host side
int lg_csr_reserve(CsrWs& ws, size_t table_entries, int n_points, cudaStream_t st) {
  LG_CHECK(ws.tab.ensure((table_entries + 16) * 4, st));
  LG_CHECK(ws.sorted.ensure(((size_t)n_points + 16) * 16, st));
  LG_CHECK(ws.sums.ensure((table_entries / SCAN_TILE + 16) * 4, st));
  return LOAM_OK;
}

static size_t csr_set_box(CsrGridD& g, const int lo[3], const int hi[3]) {
  g.x0 = (float)lo[0]; g.y0 = (float)lo[1]; g.z0 = (float)lo[2];
  g.nx = std::max(1, hi[0] - lo[0] + 1);
  g.ny = std::max(1, hi[1] - lo[1] + 1);
  g.nz = std::max(1, hi[2] - lo[2] + 1);
  g.nxp = g.nx + 2;
  return (size_t)g.nz * g.ny * g.nxp + 1;  // + the guard entry in front
}

int lg_csr_build2(CsrWs& ws, const float4* pts0, int n0, const int lo0[3], const int hi0[3], const float4* pts1, int n1, const int lo1[3],
                  const int hi1[3], cudaStream_t st, long long* launches) {
  CsrJob J;
  // grid 1's table starts on a 16-byte boundary (the kernel reads four entries at a time): the padding entries hold 0
  const size_t len0 = (csr_set_box(J.g[0], lo0, hi0) + 3) & ~(size_t)3, len1 = csr_set_box(J.g[1], lo1, hi1);
  const size_t len = len0 + len1;
  if (len > (size_t)0x7fffffff) return LOAM_ENOSPC;
  int rc = lg_csr_reserve(ws, len, n0 + n1, st);
  if (rc) return rc;
  J.pts[0] = pts0; J.pts[1] = pts1;
  J.n[0] = n0; J.n[1] = n1;
  J.tab = ws.tab.as<unsigned int>();
  J.off[0] = 0; J.off[1] = (int)len0;
  J.sorted = ws.sorted.as<float4>();
  for (int w = 0; w < 2; w++) {
    J.g[w].sorted = J.sorted;
    J.g[w].E = J.tab + J.off[w];
    J.g[w].n = J.n[w];
  }
  ws.d[0] = J.g[0];
  ws.d[1] = J.g[1];
  LgProfScope prof_scope(LGK_GRID, st, (double)(n0 + n1));
  LG_CHECK(cudaMemsetAsync(J.tab, 0, len * 4, st));
  const int pb0 = lg_div_up(n0, 256), pb1 = lg_div_up(n1, 256);
  const int nb = (int)((len + SCAN_TILE - 1) / SCAN_TILE);
  if (pb0 + pb1 > 0) {
    csr_count_kernel<<<pb0 + pb1, 256, 0, st>>>(J, pb0);
    (*launches)++;
  }
  // the guard entries hold 0, every cell entry its count: after the exclusive scan a cell entry holds the START of its run
  // in the concatenated sorted array (grid 1 continues after grid 0), the scatter turns it into the END
  scan_sums_kernel<<<nb, SCAN_NT, 0, st>>>(J.tab, len, ws.sums.as<unsigned int>());
  scan_top_kernel<<<1, SCAN_NT, 0, st>>>(ws.sums.as<unsigned int>(), nb);
  scan_apply_kernel<<<nb, SCAN_NT, 0, st>>>(J.tab, len, ws.sums.as<unsigned int>());
  (*launches) += 3;
  if (pb0 + pb1 > 0) {
    csr_fill_kernel<<<pb0 + pb1, 256, 0, st>>>(J, pb0);
    (*launches)++;
  }
  LG_CHECK(cudaGetLastError());
  return LOAM_OK;
}

int lg_csr_bbox_launch(const float4* pts0, int n0, const float4* pts1, int n1, int* d_bb12, cudaStream_t st, long long* launches) {
  const int init[12] = {INT_MAX, INT_MAX, INT_MAX, INT_MIN, INT_MIN, INT_MIN, INT_MAX, INT_MAX, INT_MAX, INT_MIN, INT_MIN, INT_MIN};
  LG_CHECK(cudaMemcpyAsync(d_bb12, init, sizeof(init), cudaMemcpyHostToDevice, st));
  LG_CHECK(cudaStreamSynchronize(st));  // `init` is on the stack
  if (n0 + n1 > 0) {
    csr_bbox_kernel<<<lg_div_up(n0 + n1, 256), 256, 0, st>>>(pts0, n0, pts1, n1, d_bb12);
    (*launches)++;
  }
  LG_CHECK(cudaGetLastError());
  return LOAM_OK;
}

static const void* gn_kernel_of(int sub) {
  // one thread per query: two CTAs per SM without spills measured faster (353 us) than three at 80 registers (387 us)
  if (sub == 1) return (const void*)map_gn_kernel<1, 2>;
  return sub == 4 ? (const void*)map_gn_kernel<4, 2> : (const void*)map_gn_kernel<8, 2>;
}

static int lg_map_gn_grid(int nq, int device, int sub, int max_ctas, int* tile_out) {
  static int per_sm[3][64] = {{0}}, sms[64] = {0};
  const int d = device & 63, m = sub == 1 ? 0 : (sub == 4 ? 1 : 2);
  if (!per_sm[m][d]) {
    int occ = 0, n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gn_kernel_of(sub), GN_NT, 0) != cudaSuccess || occ < 1) occ = 1;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || n < 1) n = 1;
    per_sm[m][d] = occ;
    sms[d] = n;
  }
  // tile: the largest of 256 / 128 / 64 / 32 stack points per CTA step that still gives every resident CTA a tile
  // small problems (one sweep's stack against the local map): one CTA per SM at most -- a cooperative grid that fills
  // the register files would lock the kernels of the other pipeline stages (extraction, odometry) out while it runs
  // loam_params.gn_max_ctas (LOAM_GN_MAX_CTAS overrides it): cap of the cooperative grid for sweep-sized stacks -- several
  // sequences sharing one GPU run their mapping loops side by side instead of queueing for all the SMs
  static const int cap_env = getenv("LOAM_GN_MAX_CTAS") ? atoi(getenv("LOAM_GN_MAX_CTAS")) : 0;
  const int cap = cap_env > 0 ? cap_env : max_ctas;
  int resident = nq <= (1 << 16) ? sms[d] : per_sm[m][d] * sms[d];
  if (cap > 0 && nq <= (1 << 16)) resident = std::min(resident, cap);
  int tile = GN_TILE;
  while (tile > 32 && (nq + tile - 1) / tile < resident) tile >>= 1;
  *tile_out = tile;
  const int ntiles = std::max(1, (nq + tile - 1) / tile);
  return std::min(ntiles, resident);
}

int lg_map_gn_launch(MapGnWs& ws, MapGnArgs& A, int device, cudaStream_t st, long long* launches) {
  const int nq = A.n_cs + A.n_ss;
  // Layout.  Groups of 8 lanes per query (the lanes scan the nine runs together: lowest latency per query) up to 2^19 stack
  // points; one thread per query with cell pruning above (measured, 1 M points against a 20 M-point map: 356 vs 496 us per
  // iteration; 200 k against 2 M, where the map sits in L2: 166 vs 147).  LOAM_GN_SUB = 1 / 4 / 8 forces a layout: all of
  // them give the same neighbours and rows, the sums up to the order of addition.
  static const int sub_env = getenv("LOAM_GN_SUB") ? atoi(getenv("LOAM_GN_SUB")) : 0;
  const int sub = sub_env == 1 || sub_env == 4 || sub_env == 8 ? sub_env : (nq > (1 << 19) ? 1 : 8);
  const int grid = lg_map_gn_grid(nq, device, sub, A.max_ctas, &A.tile);
  LG_CHECK(ws.partials.ensure((size_t)grid * 28 * 8 + 64, st));
  if (!ws.sync.p) {
    LG_CHECK(ws.sync.ensure(64 * 4 + 28 * 4 * 8, st));
    LG_CHECK(cudaMemsetAsync(ws.sync.p, 0, 64 * 4 + 28 * 4 * 8, st));
  }
  A.partials = ws.partials.as<double>();
  A.ticket = ws.sync.as<unsigned int>();
  A.gen = ws.sync.as<unsigned int>() + 1;
  A.state = ws.sync.as<float>() + 16;
  A.work = ws.sync.as<unsigned int>() + 2;
  A.fx = reinterpret_cast<unsigned long long*>(ws.sync.as<unsigned int>() + 64);
  static const bool dbg_env = getenv("LOAM_GN_DEBUG") != nullptr;
  A.dbg = nullptr;
  if (dbg_env) {
    LG_CHECK(ws.dbg.ensure(16 * 8 * 8, st));
    LG_CHECK(cudaMemsetAsync(ws.dbg.p, 0, 16 * 8 * 8, st));
    A.dbg = ws.dbg.as<unsigned long long>();
  }
  LgProfScope prof_scope(LGK_MAP_KNN, st, (double)nq);
  void* args[] = {(void*)&A};
  LG_CHECK(cudaLaunchCooperativeKernel(gn_kernel_of(sub), dim3(grid), dim3(GN_NT), args, 0, st));
  (*launches)++;
  if (dbg_env) {
    unsigned long long hb[16 * 8];
    LG_CHECK(cudaStreamSynchronize(st));
    LG_CHECK(cudaMemcpy(hb, ws.dbg.p, sizeof(hb), cudaMemcpyDeviceToHost));
    for (int i = 0; i < A.it1 - A.it0 && i < 16 && hb[i * 8]; i++) {
      const unsigned long long* t = hb + i * 8;
      auto us = [&](int a, int b) { return t[a] && t[b] ? ((double)t[b] - (double)t[a]) * 1e-3 : 0.0; };
      fprintf(stderr, "[gn] nq %d grid %d tile %d iter %d: CTA 0 search %5.1f us, fit + rows %5.1f | all CTAs done +%5.1f | totals +%4.1f | solve +%4.1f | barrier +%4.1f | iteration %6.1f\n",
              nq, grid, A.tile, A.it0 + i, us(0, 6), us(6, 1), us(1, 2), us(2, 3), us(3, 4), us(4, 5), us(0, 5));
    }
  }
  return LOAM_OK;
}
