// Host-side interface of lg_map.cu (laserMapping kernels).
#pragma once
#include <algorithm>

#include "lg_common.cuh"

struct MapT {  // transformTobeMapped[6] (LM:108) + sin/cos of its three angles evaluated on the host
  float t[6];
  SinCos3 sc;
};

struct CubeGeom {  // LM:69-75
  int W, H, D, cenW, cenH, cenD;
};

constexpr int GRID_INLINE = 7;  // points stored inside a bucket

struct alignas(128) GridBucket {  // one 128-byte line per occupied 1 m cell
  unsigned long long key;        // cell key, ~0 = empty
  int count;                     // points in the cell
  int start;                     // first position in `sorted` (only when count > GRID_INLINE)
  float4 pts[GRID_INLINE];       // {x, y, z, original index as int bits}
};

struct GridD {  // voxel hash over one map cloud (cell = 1 m), open addressing over 128-byte buckets
  GridBucket* buckets;
  int* fill;       // build-time scatter cursor per slot
  int* cursor;     // global allocation cursor into `sorted`
  int* slot_of;    // slot of every map point
  float4* sorted;  // cells with more than GRID_INLINE points: all their points, contiguous
  const unsigned int* occ;  // one bit per slot (occupied), small enough to stay in L2: empty cells cost no DRAM access
  int bits;        // log2(slots)
  int n;
};

struct GridWs {
  DevBuf keys, ints, slot_of, sorted;
  GridD d;
  void release() { keys.release(); ints.release(); slot_of.release(); sorted.release(); }
};

struct MapIterWs {
  DevBuf nbr, partials, ticket;
  void release() { nbr.release(); partials.release(); ticket.release(); }
};

int lg_map_stack_launch(const MapT& T, const float4* in0, float4* out0, int n0, const float4* in1, float4* out1, int n1, cudaStream_t st,
                        long long* launches);
int lg_map_register_launch(const MapT& T, const float4* in, float4* out, int n, cudaStream_t st, long long* launches);
int lg_grid_build(GridWs& ws, const float4* pts, int n, cudaStream_t st, long long* launches);
int lg_map_iter_launch(MapIterWs& ws, const MapT& T, const float4* corner_stack, int n_cs, const float4* surf_stack, int n_ss, const GridD& gc,
                       const GridD& gs, const float4* corner_map, const float4* surf_map, double* out28, unsigned long long seq, cudaStream_t st,
                       long long* launches);
int lg_map_insert_launch(const MapT& T, const CubeGeom& cg, const float4* corner_stack, int n_cs, const float4* surf_stack, int n_ss,
                         float4* sel_out, unsigned long long* keys, unsigned int* vals, cudaStream_t st, long long* launches);
int lg_map_runs_launch(const unsigned long long* keys, const unsigned int* vals, const float4* sel, int n, float4* sorted_sel, int* n_runs,
                       int2* runs, int cap_runs, cudaStream_t st, long long* launches);
