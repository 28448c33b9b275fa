// Shared device/host helpers for libloamgpu (sm_100a).  The whole library is compiled with -fmad=false so that every
// fp32 expression keeps the reference's evaluation order without FMA contraction.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <cstdio>

#include "../../include/loamgpu.h"

#define LG_CHECK(expr)                                                         \
  do {                                                                         \
    cudaError_t _e = (expr);                                                   \
    if (_e != cudaSuccess) {                                                   \
      lg_set_error(cudaGetErrorString(_e), __FILE__, __LINE__);                \
      return LOAM_ECUDA;                                                       \
    }                                                                          \
  } while (0)

void lg_set_error(const char* msg, const char* file, int line);

// Device buffer that only grows.
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes, cudaStream_t st, bool keep = false) {
    if (bytes <= cap) return cudaSuccess;
    size_t ncap = bytes + bytes / 2 + 256;
    void* np = nullptr;
    cudaError_t e = cudaMalloc(&np, ncap);
    if (e != cudaSuccess) return e;
    if (p) {
      if (keep) {
        e = cudaMemcpyAsync(np, p, cap, cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) return e;
      }
      cudaStreamSynchronize(st);
      cudaFree(p);
    }
    p = np;
    cap = ncap;
    return cudaSuccess;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <typename T>
  T* as() const { return (T*)p; }
};

static inline int lg_div_up(int a, int b) { return (a + b - 1) / b; }

#ifdef __CUDACC__

// squared distance in the order FLANN's L2_Simple accumulates it: ((dx*dx)+(dy*dy))+(dz*dz)
__device__ __forceinline__ float lg_sqdist(float ax, float ay, float az, float bx, float by, float bz) {
  float dx = ax - bx, dy = ay - by, dz = az - bz;
  return ((dx * dx) + (dy * dy)) + (dz * dz);
}

// (d2, idx) -> 64-bit key whose unsigned order is the kNN tie rule (d2 ascending, idx ascending); d2 >= 0.
__device__ __forceinline__ unsigned long long lg_pack_nbr(float d2, int idx) {
  return ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned int)idx;
}
__device__ __forceinline__ float lg_nbr_d2(unsigned long long k) { return __uint_as_float((unsigned int)(k >> 32)); }
__device__ __forceinline__ int lg_nbr_idx(unsigned long long k) { return (int)(unsigned int)(k & 0xffffffffull); }

// sin/cos of an fp32 angle evaluated in fp64 and rounded once: the closest a GPU gets to glibc's (almost always
// correctly rounded) sinf/cosf that the reference's `sin(float)`/`cos(float)` calls resolve to.
__device__ __forceinline__ void lg_sincosf_cr(float a, float* s, float* c) {
  double sd, cd;
  sincos((double)a, &sd, &cd);
  *s = (float)sd;
  *c = (float)cd;
}

struct SinCos3 {  // sin/cos of rx, ry, rz evaluated on the HOST with libm (bit-identical to the oracle's calls)
  float srx, crx, sry, cry, srz, crz;
};

// warp-level sum of a double
__device__ __forceinline__ double lg_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide exclusive scan of one int per thread; *total = sum.  s_warp needs NT/32 + 1 ints.  Ends with a barrier.
template <int NT>
__device__ __forceinline__ int block_excl_scan(int v, int* total, int* s_warp /* >= NT/32 + 1 */) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) s_warp[w] = inc;
  __syncthreads();
  if (w == 0) {
    int x = lane < NT / 32 ? s_warp[lane] : 0;
    int xi = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, xi, o);
      if (lane >= o) xi += t;
    }
    if (lane < NT / 32) s_warp[lane] = xi - x;
    if (lane == 31) s_warp[NT / 32] = xi;
  }
  __syncthreads();
  int r = s_warp[w] + inc - v;
  *total = s_warp[NT / 32];
  __syncthreads();
  return r;
}

#endif  // __CUDACC__
