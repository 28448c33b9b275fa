// Host-side interface of lg_voxel.cu (voxel-grid down-sampling, radix sort, table-driven gather).
#pragma once
#include <algorithm>

#include "lg_common.cuh"

struct VoxSegD {               // one independent voxel-grid problem (device-resident descriptor)
  const float4* in;            // input points
  const unsigned char* valid;  // optional per-point mask (1 = take part), may be null
  float4* out;                 // output, capacity >= n
  int* out_count;              // receives the number of occupied cells (or -1 on small-path capacity overflow)
  int n;
  float leaf;
};

struct CopyEnt {  // table-driven gather: dst[dst_off + i] = src[i], i < n
  const float4* src;
  int n;
  int dst_off;
};

struct RadixWs {
  DevBuf keysA, keysB, valsA, valsB, hist, perm;
  void release() { keysA.release(); keysB.release(); valsA.release(); valsB.release(); hist.release(); perm.release(); }
};
struct VoxBigWs {
  RadixWs rs;
  DevBuf bb, block_sums, keys_old, vals_old, keys_m, vals_m;
  void release() {
    rs.release(); bb.release(); block_sums.release(); keys_old.release(); vals_old.release(); keys_m.release(); vals_m.release();
  }
};

// One CTA per segment, shared-memory bitonic sort; max_seg_hint picks the 4096- or 16384-point instantiation.
int lg_vox_small(const VoxSegD* d_segs, int nseg, int max_seg_hint, int* d_overflow, cudaStream_t st, long long* launches);
// 4 k .. 64 k points per segment (no `valid` mask): the cell-id range is split over 32 CTAs per segment; sets *d_overflow
// when a sub-range exceeds its shared-memory capacity (caller falls back to lg_vox_small / lg_vox_big).
int lg_vox_split(DevBuf& staging, DevBuf& counts, const VoxSegD* d_segs, int nseg, int* d_overflow, cudaStream_t st, long long* launches);
// Any size: segments are contiguous in d_in, d_seg_off[nseg + 1]; outputs contiguous in d_out, per-segment
// [d_out_start[s], d_out_end[s]).
int lg_vox_big(VoxBigWs& ws, const float4* d_in, const int* d_seg_off, const float* d_seg_leaf, int nseg, int M, float4* d_out,
               int* d_out_start, int* d_out_end, cudaStream_t st, long long* launches);
// Merge path for cube-sized segments whose old cloud is already voxel-gridded (see lg_voxel.cu): d_in = [old | new].
// lg_vox_small over several segment arrays at once: grid.y picks {segment array, overflow flag} from the device tables
int lg_vox_small_batch(const VoxSegD* const* d_seg_tab, int* const* d_overflow_tab, int max_nseg, int B, int max_seg_hint, cudaStream_t st,
                       long long* launches);
int lg_vox_merge(VoxBigWs& ws, const float4* d_in, const int* d_seg_off_old, const int* d_seg_off_new, const int* d_seg_off,
                 const float* d_seg_leaf, int nseg, int n_old, int n_new, float4* d_out, int* d_out_start, int* d_out_end, int* d_flags,
                 cudaStream_t st, long long* launches);
int lg_radix_ensure(RadixWs& ws, int n, cudaStream_t st);
// Stable LSD radix sort of ws.keysA/valsA over the low `bits` bits; *result_in_b tells which buffer holds the result.
int lg_radix_sort(RadixWs& ws, int n, int bits, cudaStream_t st, long long* launches, int* result_in_b);
int lg_gather(const CopyEnt* d_ents, int nent, int max_n, float4* d_dst, cudaStream_t st, long long* launches);
