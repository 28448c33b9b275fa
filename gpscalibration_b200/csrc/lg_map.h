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

struct GridD {  // voxel hash over one map cloud (cell = 1 m)
  unsigned long long* keys;  // cell key per slot, ~0 = empty
  int* count;                // points in the cell
  int* start;                // first position in `sorted`
  int* fill;                 // scatter cursor
  int* cursor;               // global allocation cursor
  int* slot_of;              // slot of every map point
  float4* sorted;            // cell-sorted copy {x, y, z, original index as int bits}
  int bits;                  // log2(slots)
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
