// Shared device/host helpers for libloamgpu (sm_100a).  The whole library is compiled with -fmad=false so that every
// fp32 expression keeps the reference's evaluation order without FMA contraction.
#pragma once
#include <stdio.h>
#include <stdlib.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <cstdio>
#include <vector>

#include "../../include/loamgpu.h"
#include "lg_libm.cuh"

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
    size_t ncap = 2 * bytes + (1u << 20);  // grow rarely: cudaMalloc / cudaFree cost milliseconds and stall the stream
    static const bool trace = getenv("LOAM_TRACE_ALLOC") != nullptr;
    if (trace) fprintf(stderr, "[loamgpu] DevBuf %p grows %zu -> %zu bytes\n", (void*)this, cap, ncap);
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


// ---- optional per-kernel-class CUDA-event timing (bench.py's roofline object); off by default, zero cost when off
enum LgKernelClass {
  LGK_EXTRACT = 0,   // K1-K4: every scanRegistration kernel incl. the per-ring voxel grid
  LGK_ODOM_KNN = 1,  // K8: odom_knn_kernel
  LGK_ODOM_ITER = 2, // K8/K9: odom_iter_kernel
  LGK_TO_END = 3,    // K6: odom_to_end_kernel
  LGK_MAP_STACK = 4, // K6: map_stack_kernel / map_register_kernel
  LGK_VOXEL = 5,     // K5: voxel grids of stacks, cubes, surround (incl. their sorts)
  LGK_GATHER = 6,    // K11: table-driven gathers
  LGK_GRID = 7,      // K7: voxel-hash build
  LGK_MAP_KNN = 8,   // K10: map_knn_kernel
  LGK_MAP_FIT = 9,   // K10: map_fit_kernel
  LGK_INSERT = 10,   // K11: map_insert_kernel + cube sort + runs
  LGK_SR_SELECT = 11,  // K4: sr_select_kernel alone (the largest single launch of a sweep)
  LGK_COUNT = 12
};
struct LgProf {
  bool on = false;
  struct Pair { int cls; cudaEvent_t a, b; };
  std::vector<cudaEvent_t> pool;
  std::vector<Pair> pending;
  double ms[LGK_COUNT] = {0};
  double units[LGK_COUNT] = {0};
  long long scopes[LGK_COUNT] = {0};
  cudaEvent_t get() {
    if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
  }
  void resolve(cudaStream_t st) {
    if (pending.empty()) return;
    cudaStreamSynchronize(st);
    for (auto& p : pending) {
      float t = 0.f;
      cudaEventElapsedTime(&t, p.a, p.b);
      ms[p.cls] += t;
      pool.push_back(p.a);
      pool.push_back(p.b);
    }
    pending.clear();
  }
  void release() { for (auto e : pool) cudaEventDestroy(e); pool.clear(); }
};
extern thread_local LgProf* g_lg_prof;
struct LgProfScope {
  LgProf* p; int cls; cudaStream_t st; cudaEvent_t a;
  LgProfScope(int c, cudaStream_t s, double units) : p(g_lg_prof), cls(c), st(s), a(nullptr) {
    if (!p) return;
    a = p->get();
    cudaEventRecord(a, st);
    p->units[cls] += units;
    p->scopes[cls]++;
  }
  ~LgProfScope() {
    if (!p) return;
    cudaEvent_t b = p->get();
    cudaEventRecord(b, st);
    p->pending.push_back({cls, a, b});
    if (p->pending.size() > 8192) p->resolve(st);
  }
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

// sin/cos of an fp32 angle exactly as the host libm (glibc sinf / cosf) returns them — see lg_libm.cuh.
__device__ __forceinline__ void lg_sincosf_cr(float a, float* s, float* c) {
  *s = lgm_sinf(a);
  *c = lgm_cosf(a);
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
