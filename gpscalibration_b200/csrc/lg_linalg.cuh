// Small fp32 dense linear algebra used inside the mapping kernel and by the host-side Gauss-Newton step.
// The reference calls OpenCV for these (un-vendored, unpinned): cv::eigen on a symmetric 3x3 (LM:810) and 6x6
// (LO:982, LM:975), cv::solve(DECOMP_QR) on 5x3 (LM:875) and 6x6 (LO:975, LM:968), cv::Mat::inv (LO:997, LM:990) and
// Mat * Mat.  These are restatements of the published algorithms (Householder QR, max-pivot Jacobi sorted
// descending with row eigenvectors, LU with partial pivoting, GEMM with double accumulators) in a fixed fp32
// operation order, compiled without FMA contraction on both host (-ffp-contract=off) and device (-fmad=false).
#pragma once
#include <math.h>
#include <string.h>

#ifdef __CUDACC__
#define LG_HD __host__ __device__ __forceinline__
#define LG_UNROLL _Pragma("unroll")
#else
#define LG_HD inline
#define LG_UNROLL
#endif

// Least squares A(M x N) x = b by Householder reflections; columns left to right, rows top to bottom.  Every loop has
// compile-time bounds and is fully unrolled on the device, so the working arrays live in registers (a 6x6 solve on one
// GPU thread took ~10 us out of local memory and ~1 us out of registers).
template <int M, int N>
LG_HD bool lg_qr_solve(const float* A0, const float* b0, float* x) {
  float a[M * N];
  float b[M];
LG_UNROLL
  for (int i = 0; i < M * N; i++) a[i] = A0[i];
LG_UNROLL
  for (int i = 0; i < M; i++) b[i] = b0[i];
LG_UNROLL
  for (int k = 0; k < N; k++) {
    float nrm2 = 0.f;
LG_UNROLL
    for (int i = k; i < M; i++) nrm2 = nrm2 + a[i * N + k] * a[i * N + k];
    float nrm = sqrtf(nrm2);
    if (nrm == 0.f) {
LG_UNROLL
      for (int i = 0; i < N; i++) x[i] = 0.f;
      return false;
    }
    float alpha = (a[k * N + k] > 0.f) ? -nrm : nrm;
    float v[M];
LG_UNROLL
    for (int i = 0; i < M; i++) v[i] = 0.f;
    v[k] = a[k * N + k] - alpha;
LG_UNROLL
    for (int i = k + 1; i < M; i++) v[i] = a[i * N + k];
    float vn2 = 0.f;
LG_UNROLL
    for (int i = k; i < M; i++) vn2 = vn2 + v[i] * v[i];
LG_UNROLL
    for (int j = k + 1; j < N; j++) {
      float s = 0.f;
LG_UNROLL
      for (int i = k; i < M; i++) s = s + v[i] * a[i * N + j];
      float f = (2.f * s) / vn2;
LG_UNROLL
      for (int i = k; i < M; i++) a[i * N + j] = a[i * N + j] - f * v[i];
    }
    {
      float s = 0.f;
LG_UNROLL
      for (int i = k; i < M; i++) s = s + v[i] * b[i];
      float f = (2.f * s) / vn2;
LG_UNROLL
      for (int i = k; i < M; i++) b[i] = b[i] - f * v[i];
    }
    a[k * N + k] = alpha;
LG_UNROLL
    for (int i = k + 1; i < M; i++) a[i * N + k] = 0.f;
  }
LG_UNROLL
  for (int i = N - 1; i >= 0; i--) {
    float s = b[i];
LG_UNROLL
    for (int j = i + 1; j < N; j++) s = s - a[i * N + j] * x[j];
    x[i] = s / a[i * N + i];
  }
  return true;
}

#ifdef __CUDACC__
// lg_qr_solve<6, 6> spread over seven lanes of a warp: lane c < 6 owns column c of A, lane 6 owns b.  Per reflection the
// owner of the pivot column forms the Householder vector and broadcasts it, the lanes to its right apply it to their
// own column.  Every element goes through exactly the operations of the one-thread version in the same order, so the
// result is bit-identical; the six serial column updates per reflection collapse into one.  All 32 lanes must call;
// all lanes return the solution.
__device__ __forceinline__ bool lg_qr_solve6_warp(const float* A0 /* 6x6 row-major */, const float* b0, int lane, float* x) {
  float col[6];
#pragma unroll
  for (int i = 0; i < 6; i++) {
    float v = b0[i];
#pragma unroll
    for (int c = 0; c < 6; c++) v = lane == c ? A0[i * 6 + c] : v;
    col[i] = v;
  }
#pragma unroll
  for (int k = 0; k < 6; k++) {
    float v[6], vn2 = 0.f, nrm = 0.f;
#pragma unroll
    for (int i = 0; i < 6; i++) v[i] = 0.f;
    if (lane == k) {
      float nrm2 = 0.f;
#pragma unroll
      for (int i = k; i < 6; i++) nrm2 = nrm2 + col[i] * col[i];
      nrm = sqrtf(nrm2);
      const float alpha = (col[k] > 0.f) ? -nrm : nrm;
      v[k] = col[k] - alpha;
#pragma unroll
      for (int i = k + 1; i < 6; i++) v[i] = col[i];
#pragma unroll
      for (int i = k; i < 6; i++) vn2 = vn2 + v[i] * v[i];
      col[k] = alpha;
#pragma unroll
      for (int i = k + 1; i < 6; i++) col[i] = 0.f;
    }
    nrm = __shfl_sync(0xffffffffu, nrm, k);
    if (nrm == 0.f) {
#pragma unroll
      for (int i = 0; i < 6; i++) x[i] = 0.f;
      return false;
    }
    vn2 = __shfl_sync(0xffffffffu, vn2, k);
#pragma unroll
    for (int i = k; i < 6; i++) v[i] = __shfl_sync(0xffffffffu, v[i], k);
    if (lane > k && lane <= 6) {
      float s = 0.f;
#pragma unroll
      for (int i = k; i < 6; i++) s = s + v[i] * col[i];
      const float f = (2.f * s) / vn2;
#pragma unroll
      for (int i = k; i < 6; i++) col[i] = col[i] - f * v[i];
    }
  }
  // back-substitution: every lane gathers the triangle and the right-hand side and solves for itself
  float r[6][6], rb[6];
#pragma unroll
  for (int i = 0; i < 6; i++) {
    rb[i] = __shfl_sync(0xffffffffu, col[i], 6);
#pragma unroll
    for (int j = i; j < 6; j++) r[i][j] = __shfl_sync(0xffffffffu, col[i], j);
  }
#pragma unroll
  for (int i = 5; i >= 0; i--) {
    float s = rb[i];
#pragma unroll
    for (int j = i + 1; j < 6; j++) s = s - r[i][j] * x[j];
    x[i] = s / r[i][i];
  }
  return true;
}
#endif

// Symmetric eigen-decomposition.  W descending, eigenvectors are the ROWS of V.  Pivot = first largest |a_kl| of the
// upper triangle (row-major scan); stops at |pivot| <= FLT_EPSILON or after 30 N^2 rotations.
template <int N>
LG_HD void lg_jacobi_eigen(const float* A0, float* W, float* V) {
  float A[N * N];
  for (int i = 0; i < N * N; i++) A[i] = A0[i];
  for (int i = 0; i < N; i++)
    for (int j = 0; j < N; j++) V[i * N + j] = (i == j) ? 1.f : 0.f;
  for (int k = 0; k < N; k++) W[k] = A[k * N + k];
  const float eps = 1.1920929e-07f;
  const int maxIters = N * N * 30;
  for (int it = 0; it < maxIters; it++) {
    int k = 0, l = 1;
    float mv = fabsf(A[0 * N + 1]);
    for (int i = 0; i < N - 1; i++)
      for (int j = i + 1; j < N; j++) {
        float val = fabsf(A[i * N + j]);
        if (mv < val) {
          mv = val;
          k = i;
          l = j;
        }
      }
    float p = A[k * N + l];
    if (fabsf(p) <= eps) break;
    float y = (W[l] - W[k]) * 0.5f;
    float t = fabsf(y) + sqrtf(p * p + y * y);
    float s = sqrtf(p * p + t * t);
    float c = t / s;
    s = p / s;
    t = (p / t) * p;
    if (y < 0.f) {
      s = -s;
      t = -t;
    }
    A[k * N + l] = 0.f;
    W[k] = W[k] - t;
    W[l] = W[l] + t;
    float a0, b0;
#define LG_ROT(v0, v1) a0 = (v0), b0 = (v1), (v0) = a0 * c - b0 * s, (v1) = a0 * s + b0 * c
    for (int i = 0; i < k; i++) LG_ROT(A[i * N + k], A[i * N + l]);
    for (int i = k + 1; i < l; i++) LG_ROT(A[k * N + i], A[i * N + l]);
    for (int i = l + 1; i < N; i++) LG_ROT(A[k * N + i], A[l * N + i]);
    for (int i = 0; i < N; i++) LG_ROT(V[k * N + i], V[l * N + i]);
#undef LG_ROT
  }
  for (int k = 0; k < N - 1; k++) {
    int m = k;
    for (int i = k + 1; i < N; i++)
      if (W[m] < W[i]) m = i;
    if (k != m) {
      float tw = W[m];
      W[m] = W[k];
      W[k] = tw;
      for (int i = 0; i < N; i++) {
        float tv = V[m * N + i];
        V[m * N + i] = V[k * N + i];
        V[k * N + i] = tv;
      }
    }
  }
}

#ifdef __CUDACC__
// lg_jacobi_eigen<3> with every index static (the generic version indexes A, W, V with the pivot position, which puts them
// into local memory on the device): the three possible pivots are three copies of the same rotation.  Same operations in
// the same order, element for element -- bit-identical results.  Only the upper triangle of A is ever read.
__device__ __forceinline__ void lg_jacobi_eigen3(const float* A0, float* W, float* V) {
  float a01 = A0[1], a02 = A0[2], a12 = A0[5];
  float w0 = A0[0], w1 = A0[4], w2 = A0[8];
  float v00 = 1.f, v01 = 0.f, v02 = 0.f, v10 = 0.f, v11 = 1.f, v12 = 0.f, v20 = 0.f, v21 = 0.f, v22 = 1.f;
  const float eps = 1.1920929e-07f;
#define LG_ROT(v0, v1) a0 = (v0), b0 = (v1), (v0) = a0 * c - b0 * s, (v1) = a0 * s + b0 * c
#define LG_JAC3(P, WK, WL, R0, R1, VK0, VK1, VK2, VL0, VL1, VL2)   \
  {                                                                \
    const float p = P;                                             \
    const float y = (WL - WK) * 0.5f;                              \
    float t = fabsf(y) + sqrtf(p * p + y * y);                     \
    float s = sqrtf(p * p + t * t);                                \
    const float c = t / s;                                         \
    s = p / s;                                                     \
    t = (p / t) * p;                                               \
    if (y < 0.f) {                                                 \
      s = -s;                                                      \
      t = -t;                                                      \
    }                                                              \
    P = 0.f;                                                       \
    WK = WK - t;                                                   \
    WL = WL + t;                                                   \
    float a0, b0;                                                  \
    LG_ROT(R0, R1);                                                \
    LG_ROT(VK0, VL0);                                              \
    LG_ROT(VK1, VL1);                                              \
    LG_ROT(VK2, VL2);                                              \
  }
#pragma unroll 1
  for (int it = 0; it < 3 * 3 * 30; it++) {
    int piv = 0;
    float mv = fabsf(a01);
    if (mv < fabsf(a02)) {
      mv = fabsf(a02);
      piv = 1;
    }
    if (mv < fabsf(a12)) {
      mv = fabsf(a12);
      piv = 2;
    }
    if (mv <= eps) break;  // |pivot| <= eps
    if (piv == 0) {         // (k, l) = (0, 1): i > l rotates (A[0][2], A[1][2])
      LG_JAC3(a01, w0, w1, a02, a12, v00, v01, v02, v10, v11, v12)
    } else if (piv == 1) {  // (0, 2): k < i < l rotates (A[0][1], A[1][2])
      LG_JAC3(a02, w0, w2, a01, a12, v00, v01, v02, v20, v21, v22)
    } else {                // (1, 2): i < k rotates (A[0][1], A[0][2])
      LG_JAC3(a12, w1, w2, a01, a02, v10, v11, v12, v20, v21, v22)
    }
  }
#undef LG_JAC3
#undef LG_ROT
#define LG_SWAPF(x, y) { const float t_ = (x); (x) = (y); (y) = t_; }
  {  // selection sort, descending, first maximum wins (k = 0)
    int m = 0;
    float wm = w0;
    if (wm < w1) { m = 1; wm = w1; }
    if (wm < w2) { m = 2; }
    if (m == 1) { LG_SWAPF(w0, w1) LG_SWAPF(v00, v10) LG_SWAPF(v01, v11) LG_SWAPF(v02, v12) }
    if (m == 2) { LG_SWAPF(w0, w2) LG_SWAPF(v00, v20) LG_SWAPF(v01, v21) LG_SWAPF(v02, v22) }
  }
  if (w1 < w2) { LG_SWAPF(w1, w2) LG_SWAPF(v10, v20) LG_SWAPF(v11, v21) LG_SWAPF(v12, v22) }
#undef LG_SWAPF
  W[0] = w0; W[1] = w1; W[2] = w2;
  V[0] = v00; V[1] = v01; V[2] = v02; V[3] = v10; V[4] = v11; V[5] = v12; V[6] = v20; V[7] = v21; V[8] = v22;
}
#endif

// ---- the pieces of the Gauss-Newton update the reference keeps on the CPU ------------------------------------------
#ifdef __CUDACC__
__host__ __device__
#endif
static inline void lg_gemm_dacc(const float* A, const float* B, float* C, int m, int k, int n) {
  for (int i = 0; i < m; i++)
    for (int j = 0; j < n; j++) {
      double s = 0.0;
      for (int t = 0; t < k; t++) s += (double)A[i * k + t] * (double)B[t * n + j];
      C[i * n + j] = (float)s;
    }
}

static inline bool lg_lu_inverse6(const float* A0, float* out) {
  const int n = 6;
  float a[36], b[36];
  for (int i = 0; i < 36; i++) a[i] = A0[i];
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) b[i * n + j] = (i == j) ? 1.f : 0.f;
  for (int i = 0; i < n; i++) {
    int k = i;
    for (int j = i + 1; j < n; j++)
      if (fabsf(a[j * n + i]) > fabsf(a[k * n + i])) k = j;
    if (fabsf(a[k * n + i]) < 1.1920929e-07f * 100.f) {
      for (int t = 0; t < 36; t++) out[t] = 0.f;
      return false;
    }
    if (k != i) {
      for (int j = i; j < n; j++) { float t = a[i * n + j]; a[i * n + j] = a[k * n + j]; a[k * n + j] = t; }
      for (int j = 0; j < n; j++) { float t = b[i * n + j]; b[i * n + j] = b[k * n + j]; b[k * n + j] = t; }
    }
    float d = -1.f / a[i * n + i];
    for (int j = i + 1; j < n; j++) {
      float alpha = a[j * n + i] * d;
      for (int t = i + 1; t < n; t++) a[j * n + t] = a[j * n + t] + alpha * a[i * n + t];
      for (int t = 0; t < n; t++) b[j * n + t] = b[j * n + t] + alpha * b[i * n + t];
    }
  }
  for (int i = n - 1; i >= 0; i--)
    for (int j = 0; j < n; j++) {
      float s = b[i * n + j];
      for (int k = i + 1; k < n; k++) s = s - a[i * n + k] * out[k * n + j];
      out[i * n + j] = s / a[i * n + i];
    }
  return true;
}

struct LgGNState {  // matP / isDegenerate live outside the iteration and sweep loops (LO:489-492, LM:399-400)
  float matP[36];
  bool degenerate;
  LgGNState() : degenerate(false) { memset(matP, 0, sizeof(matP)); }
};

// LO:975-1004 / LM:968-997
static inline void lg_gn_solve_step(const float* AtA, const float* AtB, int iter, float eig_thre, LgGNState& st, float* X) {
  lg_qr_solve<6, 6>(AtA, AtB, X);
  if (iter == 0) {
    float E[6], Vm[36], V2[36], Vinv[36];
    lg_jacobi_eigen<6>(AtA, E, Vm);
    memcpy(V2, Vm, sizeof(V2));
    st.degenerate = false;
    for (int i = 5; i >= 0; i--) {
      if (E[i] < eig_thre) {
        for (int j = 0; j < 6; j++) V2[i * 6 + j] = 0.f;
        st.degenerate = true;
      } else {
        break;
      }
    }
    lg_lu_inverse6(Vm, Vinv);
    lg_gemm_dacc(Vinv, V2, st.matP, 6, 6, 6);
  }
  if (st.degenerate) {
    float X2[6];
    for (int i = 0; i < 6; i++) X2[i] = X[i];
    lg_gemm_dacc(st.matP, X2, X, 6, 6, 1);
  }
}
