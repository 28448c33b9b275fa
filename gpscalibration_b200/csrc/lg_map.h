// Host-side interface of lg_map.cu (laserMapping kernels).
#pragma once
#include <algorithm>

#include "lg_common.cuh"
#include "lg_reduce.cuh"

struct MapT {  // transformTobeMapped[6] (LM:108) + sin/cos of its three angles evaluated on the host
  float t[6];
  SinCos3 sc;
};

struct CubeGeom {  // LM:69-75
  int W, H, D, cenW, cenH, cenD;
};

int lg_map_stack_launch(const MapT& T, const float4* in0, float4* out0, int n0, const float4* in1, float4* out1, int n1, cudaStream_t st,
                        long long* launches);
int lg_map_register_launch(const MapT& T, const float4* in, float4* out, int n, cudaStream_t st, long long* launches);
int lg_map_insert_launch(const MapT& T, const CubeGeom& cg, const float4* corner_stack, int n_cs, const float4* surf_stack, int n_ss,
                         float4* sel_out, unsigned long long* keys, unsigned int* vals, cudaStream_t st, long long* launches);
int lg_map_runs_launch(const unsigned long long* keys, const unsigned int* vals, const float4* sel, int n, float4* sorted_sel, int* n_runs,
                       int2* runs, int cap_runs, cudaStream_t st, long long* launches);

// ---- cell-sorted (CSR) grid over one map cloud + the fused Gauss-Newton kernel (lg_mapgn.cu) -------------------------------
// 1 m cells of a dense box, x fastest (rows padded by one empty cell on either side), then y, then z.  E[0] is a guard,
// E[1 + padded cell] the END of the cell's run in `sorted` (points in cell order, original index in .w): the three cells
// x-1 .. x+1 of row (y, z) are the contiguous run [E[row + cx], E[row + cx + 3]).
struct CsrGridD {
  const float4* sorted;
  const unsigned int* E;
  float x0, y0, z0;  // box origin in cells (integer-valued)
  int nx, ny, nz, nxp;
  int n;
};
struct CsrWs {
  DevBuf tab, sorted, sums, bb;
  CsrGridD d[2];  // [0] corner map, [1] surf map (built together)
  void release() { tab.release(); sorted.release(); sums.release(); bb.release(); }
};
struct MapGnWs {
  DevBuf partials, sync, nbr, dbg;
  void release() { partials.release(); sync.release(); nbr.release(); dbg.release(); }
};
struct MapGnArgs {
  float T[6], sc[6];       // transformTobeMapped (LM:108) and sin / cos of its angles (host libm) at entry
  float matP[36];          // LM:399-400 state at entry
  int degenerate;
  int it0, it1;            // iterations [it0, it1)
  int solve;               // 0: one pass, publish the 28 sums (stage-level calls); 1: Gauss-Newton loop on the device
  const float4* cstack;
  int n_cs;
  const float4* sstack;
  int n_ss;
  CsrGridD gc, gs;
  int max_ctas;            // > 0: cap of the cooperative grid for sweep-sized stacks (loam_params.gn_max_ctas)
  float slab_lo, slab_hi;  // owner rule of a sharded map: map-frame x in [lo, hi); (-inf, +inf) otherwise
  int* nbr;                // [n_cs + n_ss][5] pointSearchInd of the last iteration (optional)
  double* out;             // mailbox / device buffer, 64 doubles: [0..27] sums (solve: of iteration 0), [28] peer timeout,
                           // [31] sequence word, [32..37] final pose, [38] last iteration, [39] converged, [40] iteration 0 solved
  unsigned long long seq;
  PeerXchg px;             // world <= 1: no exchange; xseq = sequence number of this launch's FIRST iteration
  // filled by lg_map_gn_launch
  int tile;                // stack points per CTA step (32 .. 256)
  double* partials;
  unsigned int* ticket;
  unsigned int* gen;
  float* state;
  unsigned int* work;       // one thread per query: next chunk of 32 stack points (chunks are handed out dynamically)
  unsigned long long* fx;   // one thread per query: [28][4] fixed-point totals of the iteration (order-independent sums)
  unsigned long long* dbg;  // LOAM_GN_DEBUG: phase stamps [iteration][8], else null
};
int lg_csr_reserve(CsrWs& ws, size_t table_entries, int n_points, cudaStream_t st);
// Box corners in cells (inclusive) per grid; points outside are clamped into the boundary cells (still exact).
int lg_csr_build2(CsrWs& ws, const float4* pts0, int n0, const int lo0[3], const int hi0[3], const float4* pts1, int n1, const int lo1[3],
                  const int hi1[3], cudaStream_t st, long long* launches);
// d_bb12 <- {min cell x y z, max cell x y z} of both clouds (stage-level API: the caller reads it back)
int lg_csr_bbox_launch(const float4* pts0, int n0, const float4* pts1, int n1, int* d_bb12, cudaStream_t st, long long* launches);
int lg_map_gn_launch(MapGnWs& ws, MapGnArgs& A, int device, cudaStream_t st, long long* launches);
void lg_empty_launch(cudaStream_t st);
