// The 21 + 6 (+ count) term normal-equation reduction shared by odometry (LO:972-974) and mapping (LM:965-967).
// The reference forms A (n x 6, fp32), then AtA = At * A and AtB = At * B through OpenCV's GEMM, which accumulates
// float products in DOUBLE.  A float x float product is exact in double, so summing the exact products in double in
// any order reproduces the reference's sums to ~1e-16 relative; the result is rounded to fp32 once on the host.
// Layout of the 28 doubles: upper triangle of AtA row-major (00 01 .. 05 11 12 .. 55), then AtB[0..5], then n_sel.
#pragma once
#include "lg_common.cuh"

// One-shot all-reduce of the 28 sums over NVLink peer memory (sharded map): every rank owns an exchange buffer
// {slots[2 parities][LG_MAX_PEERS][32 doubles], flags[LG_MAX_PEERS]} that its peers map through CUDA IPC.
constexpr int LG_MAX_PEERS = 8;
constexpr int LG_XCHG_FLAG_OFFSET = 2 * LG_MAX_PEERS * 32;  // in doubles
constexpr int LG_XCHG_BYTES = (LG_XCHG_FLAG_OFFSET + LG_MAX_PEERS) * 8;
struct PeerXchg {
  double* buf[LG_MAX_PEERS];  // buf[r] = rank r's exchange buffer as seen from this GPU (buf[rank] is local)
  int world, rank;
  unsigned long long xseq;  // same on every rank: the iteration number since connect
  int* timeout;             // set when a peer never showed up
};

#ifdef __CUDACC__

struct Acc28 {
  double v[28];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int i = 0; i < 28; i++) v[i] = 0.0;
  }
  __device__ __forceinline__ void add_row(const float* a, float b) {
    int t = 0;
#pragma unroll
    for (int i = 0; i < 6; i++)
#pragma unroll
      for (int j = i; j < 6; j++) v[t++] += (double)a[i] * (double)a[j];
#pragma unroll
    for (int i = 0; i < 6; i++) v[21 + i] += (double)a[i] * (double)b;
    v[27] += 1.0;
  }
};

// All 28 sums of a warp with 31 shuffles instead of 140: at every step the two halves of a lane pair swap one half of
// their vector and add the other, so the vector halves as the lanes pair up; lane L (< 28) ends with the total of v[L].
__device__ __forceinline__ double lg_warp_reduce28(const double (&v)[28], int lane) {
  double a[16], b[8], c[4], d[2];
#pragma unroll
  for (int i = 0; i < 16; i++) {
    const double lo = v[i], hi = (i + 16 < 28) ? v[i + 16] : 0.0;
    const bool up = lane & 16;
    a[i] = (up ? hi : lo) + __shfl_xor_sync(0xffffffffu, up ? lo : hi, 16);
  }
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const bool up = lane & 8;
    b[i] = (up ? a[i + 8] : a[i]) + __shfl_xor_sync(0xffffffffu, up ? a[i] : a[i + 8], 8);
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const bool up = lane & 4;
    c[i] = (up ? b[i + 4] : b[i]) + __shfl_xor_sync(0xffffffffu, up ? b[i] : b[i + 4], 4);
  }
#pragma unroll
  for (int i = 0; i < 2; i++) {
    const bool up = lane & 2;
    d[i] = (up ? c[i + 2] : c[i]) + __shfl_xor_sync(0xffffffffu, up ? c[i] : c[i + 2], 2);
  }
  const bool up = lane & 1;
  return (up ? d[1] : d[0]) + __shfl_xor_sync(0xffffffffu, up ? d[0] : d[1], 1);
}

// Warp shuffle -> shared memory -> per-CTA partial in global memory; the last CTA to arrive (ticket) adds the partials
// in CTA order (deterministic) and writes the 28 results to out28 (device memory or mapped pinned host memory).
template <int NT>
__device__ __forceinline__ void lg_reduce28(Acc28& acc, double* __restrict__ partials, unsigned int* __restrict__ ticket,
                                            double* __restrict__ out28, unsigned long long seq = 0ull, const PeerXchg* px = nullptr) {
  __shared__ double s_part[NT / 32][28];
  __shared__ bool s_last;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const double wsum = lg_warp_reduce28(acc.v, lane);
  if (lane < 28) s_part[w][lane] = wsum;
  __syncthreads();
  if (tid < 28) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < NT / 32; k++) s += s_part[k][tid];
    partials[(size_t)blockIdx.x * 28 + tid] = s;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    unsigned int t = atomicAdd(ticket, 1u);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    // the per-CTA partials, added by all warps: warp w takes CTAs w, w + NT/32, ... (four loads in flight), the warps'
    // sums are then added in warp order -- a fixed order, so the result does not depend on which CTA came last
    double s = 0.0;
    if (lane < 28) {
      constexpr unsigned int W = NT / 32;
      for (unsigned int b = w; b < gridDim.x; b += 4 * W) {
        double v[4];
#pragma unroll
        for (unsigned int u = 0; u < 4; u++) v[u] = b + u * W < gridDim.x ? __ldcg(&partials[(size_t)(b + u * W) * 28 + lane]) : 0.0;
#pragma unroll
        for (unsigned int u = 0; u < 4; u++) s += v[u];
      }
    }
    __syncthreads();  // s_part is free again: every warp has passed the barrier after its first use
    if (lane < 28) s_part[w][lane] = s;
    __syncthreads();
    s = 0.0;
    if (tid < 28) {
#pragma unroll
      for (int k = 0; k < NT / 32; k++) s += s_part[k][tid];
    }
    if (tid == 0) *ticket = 0u;
    if (px != nullptr && px->world > 1) {
      // Fused all-reduce: the last CTA stores this rank's 28 sums straight into every peer's exchange buffer (NVLink
      // P2P stores), raises its flag there, waits for the peers' flags in its own buffer and adds the partials in rank
      // order -- every rank ends with bit-identical totals, one NVLink round trip after the slowest rank's reduction.
      // Slots alternate with the iteration parity: a rank one iteration ahead never overwrites sums still being read.
      const int W = px->world, me = px->rank;
      const size_t slot = ((size_t)(px->xseq & 1ull) * LG_MAX_PEERS + me) * 32;
      if (tid < 28)
        for (int r = 0; r < W; r++) px->buf[r][slot + tid] = s;
      __threadfence_system();
      __syncthreads();
      if (tid < W) *((volatile unsigned long long*)(px->buf[tid] + LG_XCHG_FLAG_OFFSET) + me) = px->xseq;
      if (tid < W) {
        volatile unsigned long long* f = (volatile unsigned long long*)(px->buf[me] + LG_XCHG_FLAG_OFFSET) + tid;
        const long long t0 = clock64();
        while (*f < px->xseq) {
          if (clock64() - t0 > (1ll << 32)) {  // ~2 s: a peer died or never called; the host turns this into an error
            *px->timeout = 1;
            break;
          }
        }
      }
      __threadfence_system();
      __syncthreads();
      if (tid < 28) {
        s = 0.0;
        const volatile double* mine = px->buf[me] + (size_t)(px->xseq & 1ull) * LG_MAX_PEERS * 32;
        for (int r = 0; r < W; r++) s += mine[(size_t)r * 32 + tid];
      }
    }
    if (tid < 28) out28[tid] = s;
    if (seq != 0ull) {  // host mailbox: publish the sequence number after the 28 values are visible system-wide
      __threadfence_system();
      __syncthreads();
      if (tid == 0) {
        *((volatile unsigned long long*)(out28 + 31)) = seq;
        __threadfence_system();
      }
    }
  }
}

#endif  // __CUDACC__

// 28 doubles -> AtA (6x6 row-major float), AtB (6 float), n_sel
#ifdef __CUDACC__
__host__ __device__
#endif
static inline void lg_unpack28(const double* r, float* AtA, float* AtB, int* n_sel) {
  int t = 0;
  for (int i = 0; i < 6; i++)
    for (int j = i; j < 6; j++) {
      float v = (float)r[t++];
      AtA[i * 6 + j] = v;
      AtA[j * 6 + i] = v;
    }
  for (int i = 0; i < 6; i++) AtB[i] = (float)r[21 + i];
  *n_sel = (int)(r[27] + 0.5);
}
