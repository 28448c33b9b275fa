// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the shipped product path;
// only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
//
// Small dense linear algebra restating the OpenCV 3 entry points the reference's hot path calls.
// OpenCV is an un-vendored third-party dependency of the reference (version unpinned,
// install/install_u1604_basic.sh:14), so these are restatements of its *published algorithms*:
//   cv::Mat(float) * cv::Mat(float)  (LO:973-974, LM:966-967)  -> gemm_f32_dacc  (GEMM, double accumulators)
//   cv::solve(..., DECOMP_QR)        (LO:975, LM:875, LM:968)  -> qr_solve       (Householder QR / least squares)
//   cv::eigen (symmetric)            (LO:982, LM:810, LM:975)  -> jacobi_eigen   (cyclic max-pivot Jacobi, descending, row vectors)
//   cv::Mat::inv() (DECOMP_LU)       (LO:997, LM:990)          -> lu_inverse     (partial pivoting)
// Cross-checked against cv2 4.13 in tests/test_oracle_thirdparty.py.
// Everything is fp32 with a fixed operation order and must be compiled with -ffp-contract=off.
#pragma once
#include <cmath>
#include <cstring>
#include <utility>

namespace orc {

// C(m x n) = A(m x k) * B(k x n); float storage, double accumulation in ascending k, one rounding at the end.
inline void gemm_f32_dacc(const float* A, const float* B, float* C, int m, int k, int n) {
  for (int i = 0; i < m; i++)
    for (int j = 0; j < n; j++) {
      double s = 0.0;
      for (int t = 0; t < k; t++) s += (double)A[i * k + t] * (double)B[t * n + j];
      C[i * n + j] = (float)s;
    }
}

// Least-squares solve of A(m x n) x = b, m >= n <= 6, by Householder reflections.  Returns false (x = 0) if a
// column norm vanishes.  Row-major A.  Fixed order: columns left to right, rows top to bottom.
inline bool qr_solve(const float* A0, const float* b0, float* x, int m, int n) {
  float a[36];
  float b[6];
  for (int i = 0; i < m * n; i++) a[i] = A0[i];
  for (int i = 0; i < m; i++) b[i] = b0[i];
  for (int k = 0; k < n; k++) {
    float nrm2 = 0.f;
    for (int i = k; i < m; i++) nrm2 = nrm2 + a[i * n + k] * a[i * n + k];
    float nrm = sqrtf(nrm2);
    if (nrm == 0.f) {
      for (int i = 0; i < n; i++) x[i] = 0.f;
      return false;
    }
    float alpha = (a[k * n + k] > 0.f) ? -nrm : nrm;
    float v[6];
    for (int i = 0; i < m; i++) v[i] = 0.f;
    v[k] = a[k * n + k] - alpha;
    for (int i = k + 1; i < m; i++) v[i] = a[i * n + k];
    float vn2 = 0.f;
    for (int i = k; i < m; i++) vn2 = vn2 + v[i] * v[i];
    for (int j = k + 1; j < n; j++) {
      float s = 0.f;
      for (int i = k; i < m; i++) s = s + v[i] * a[i * n + j];
      float f = (2.f * s) / vn2;
      for (int i = k; i < m; i++) a[i * n + j] = a[i * n + j] - f * v[i];
    }
    {
      float s = 0.f;
      for (int i = k; i < m; i++) s = s + v[i] * b[i];
      float f = (2.f * s) / vn2;
      for (int i = k; i < m; i++) b[i] = b[i] - f * v[i];
    }
    a[k * n + k] = alpha;
    for (int i = k + 1; i < m; i++) a[i * n + k] = 0.f;
  }
  for (int i = n - 1; i >= 0; i--) {
    float s = b[i];
    for (int j = i + 1; j < n; j++) s = s - a[i * n + j] * x[j];
    x[i] = s / a[i * n + i];
  }
  return true;
}

// Symmetric eigen-decomposition, n <= 6.  W: eigenvalues descending.  V: eigenvectors as ROWS (V[k*n + i]).
// Classical Jacobi: pivot = first largest |a_kl| of the upper triangle in row-major scan order; stops when the pivot
// is <= FLT_EPSILON in absolute value (OpenCV's criterion) or after 30 n^2 rotations.
inline void jacobi_eigen(const float* A0, float* W, float* V, int n) {
  float A[36];
  for (int i = 0; i < n * n; i++) A[i] = A0[i];
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) V[i * n + j] = (i == j) ? 1.f : 0.f;
  for (int k = 0; k < n; k++) W[k] = A[k * n + k];
  const float eps = 1.1920929e-07f;
  int maxIters = n * n * 30;
  if (n > 1)
    for (int it = 0; it < maxIters; it++) {
      int k = 0, l = 1;
      float mv = fabsf(A[0 * n + 1]);
      for (int i = 0; i < n - 1; i++)
        for (int j = i + 1; j < n; j++) {
          float val = fabsf(A[i * n + j]);
          if (mv < val) { mv = val; k = i; l = j; }
        }
      float p = A[k * n + l];
      if (fabsf(p) <= eps) break;
      float y = (W[l] - W[k]) * 0.5f;
      float t = fabsf(y) + sqrtf(p * p + y * y);
      float s = sqrtf(p * p + t * t);
      float c = t / s;
      s = p / s;
      t = (p / t) * p;
      if (y < 0.f) { s = -s; t = -t; }
      A[k * n + l] = 0.f;
      W[k] = W[k] - t;
      W[l] = W[l] + t;
      float a0, b0;
#define ORC_ROT(v0, v1) a0 = (v0), b0 = (v1), (v0) = a0 * c - b0 * s, (v1) = a0 * s + b0 * c
      for (int i = 0; i < k; i++) ORC_ROT(A[i * n + k], A[i * n + l]);
      for (int i = k + 1; i < l; i++) ORC_ROT(A[k * n + i], A[i * n + l]);
      for (int i = l + 1; i < n; i++) ORC_ROT(A[k * n + i], A[l * n + i]);
      for (int i = 0; i < n; i++) ORC_ROT(V[k * n + i], V[l * n + i]);
#undef ORC_ROT
    }
  for (int k = 0; k < n - 1; k++) {
    int m = k;
    for (int i = k + 1; i < n; i++)
      if (W[m] < W[i]) m = i;
    if (k != m) {
      float tw = W[m]; W[m] = W[k]; W[k] = tw;
      for (int i = 0; i < n; i++) { float tv = V[m * n + i]; V[m * n + i] = V[k * n + i]; V[k * n + i] = tv; }
    }
  }
}

// Inverse by Gaussian elimination with partial pivoting (fp32).  Returns false if singular (out = 0).
inline bool lu_inverse(const float* A0, float* out, int n) {
  float a[36], b[36];
  for (int i = 0; i < n * n; i++) a[i] = A0[i];
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) b[i * n + j] = (i == j) ? 1.f : 0.f;
  for (int i = 0; i < n; i++) {
    int k = i;
    for (int j = i + 1; j < n; j++)
      if (fabsf(a[j * n + i]) > fabsf(a[k * n + i])) k = j;
    if (fabsf(a[k * n + i]) < 1.1920929e-07f * 100.f) {
      for (int t = 0; t < n * n; t++) out[t] = 0.f;
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

// The Gauss-Newton update step both LO:972-1004 and LM:965-997 perform on the 6x6 normal equations.
// State carried across iterations AND sweeps (matP, isDegenerate are declared outside the loops, LO:489-492, LM:399-400).
struct GNState {
  float matP[36];
  bool degenerate;
  GNState() : degenerate(false) { std::memset(matP, 0, sizeof(matP)); }
};

// AtA (6x6 float), AtB (6) -> X (6).  iter==0 recomputes the degeneracy projection with threshold `eig_thre`.
inline void gn_solve_step(const float* AtA, const float* AtB, int iter, float eig_thre, GNState& st, float* X) {
  qr_solve(AtA, AtB, X, 6, 6);
  if (iter == 0) {
    float E[6], Vm[36], V2[36], Vinv[36];
    jacobi_eigen(AtA, E, Vm, 6);
    std::memcpy(V2, Vm, sizeof(V2));
    st.degenerate = false;
    for (int i = 5; i >= 0; i--) {
      if (E[i] < eig_thre) {
        for (int j = 0; j < 6; j++) V2[i * 6 + j] = 0.f;
        st.degenerate = true;
      } else {
        break;
      }
    }
    lu_inverse(Vm, Vinv, 6);
    gemm_f32_dacc(Vinv, V2, st.matP, 6, 6, 6);
  }
  if (st.degenerate) {
    float X2[6];
    for (int i = 0; i < 6; i++) X2[i] = X[i];
    gemm_f32_dacc(st.matP, X2, X, 6, 6, 1);
  }
}

// Eigen::JacobiSVD<MatrixXd>(H, ComputeThinU | ComputeThinV) of the 3x3 cross-covariance (TC:297, TC:506; Eigen 3 is an
// un-vendored dependency, CMakeLists.txt:17): restated as the two-sided Jacobi iteration it is published as -- every
// off-diagonal pair (p, q) is annihilated by a left rotation that symmetrises the 2x2 block followed by the symmetric
// Jacobi rotation -- in fp64, row-major.  H = U diag(S) V^T, S >= 0 descending.  Rows / columns that are exactly zero are
// left alone (the track problem is planar: H(2,:) = H(:,2) = 0, so U(:,2) = V(:,2) = e3).
inline void svd3_jacobi(const double* H, double* U, double* S, double* V) {
  double W[9];
  for (int i = 0; i < 9; i++) {
    W[i] = H[i];
    U[i] = V[i] = (i % 4 == 0) ? 1.0 : 0.0;
  }
  auto mul = [](const double* X, const double* Y, double* Z) {  // Z = X * Y (ascending k), Z may not alias
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        double s = X[i * 3 + 0] * Y[0 * 3 + j];
        s += X[i * 3 + 1] * Y[1 * 3 + j];
        s += X[i * 3 + 2] * Y[2 * 3 + j];
        Z[i * 3 + j] = s;
      }
  };
  const double tiny = 2.2250738585072014e-308, eps = 2.220446049250313e-16;
  for (int sweep = 0; sweep < 60; sweep++) {
    bool any = false;
    for (int q = 1; q < 3; q++)
      for (int p = 0; p < q; p++) {
        double dmax = std::fmax(std::fabs(W[0]), std::fmax(std::fabs(W[4]), std::fabs(W[8])));
        double thr = std::fmax(tiny, eps * dmax);
        if (!(std::fabs(W[p * 3 + q]) > thr || std::fabs(W[q * 3 + p]) > thr)) continue;
        any = true;
        double a = W[p * 3 + p], b = W[p * 3 + q], c = W[q * 3 + p], d = W[q * 3 + q];
        double c1 = 1.0, s1 = 0.0;
        double t = a + d, dd = c - b;
        if (std::fabs(dd) >= tiny) {
          double u = t / dd, tmp = std::sqrt(1.0 + u * u);
          s1 = 1.0 / tmp;
          c1 = u / tmp;
        }
        double x = c1 * a + s1 * c, y = c1 * b + s1 * d, z = -s1 * b + c1 * d;
        double cj = 1.0, sj = 0.0;
        if (std::fabs(y) >= tiny) {
          double tau = (z - x) / (2.0 * y);
          double tt = (tau >= 0.0 ? 1.0 : -1.0) / (std::fabs(tau) + std::sqrt(1.0 + tau * tau));
          cj = 1.0 / std::sqrt(1.0 + tt * tt);
          sj = tt * cj;
        }
        double Gt[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, J[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        Gt[p * 3 + p] = c1; Gt[q * 3 + p] = s1; Gt[p * 3 + q] = -s1; Gt[q * 3 + q] = c1;  // transpose of [[c1, s1], [-s1, c1]]
        J[p * 3 + p] = cj; J[p * 3 + q] = sj; J[q * 3 + p] = -sj; J[q * 3 + q] = cj;
        double L[9], Lt[9], T1[9], T2[9];
        mul(Gt, J, L);
        for (int i = 0; i < 3; i++)
          for (int j = 0; j < 3; j++) Lt[i * 3 + j] = L[j * 3 + i];
        mul(Lt, W, T1);
        mul(T1, J, T2);
        for (int i = 0; i < 9; i++) W[i] = T2[i];
        W[p * 3 + q] = 0.0;
        W[q * 3 + p] = 0.0;
        mul(U, L, T1);
        for (int i = 0; i < 9; i++) U[i] = T1[i];
        mul(V, J, T1);
        for (int i = 0; i < 9; i++) V[i] = T1[i];
      }
    if (!any) break;
  }
  for (int i = 0; i < 3; i++) {
    S[i] = std::fabs(W[i * 4]);
    if (W[i * 4] < 0.0)
      for (int r = 0; r < 3; r++) U[r * 3 + i] = -U[r * 3 + i];
  }
  for (int i = 0; i < 2; i++) {
    int best = i;
    for (int j = i + 1; j < 3; j++)
      if (S[j] > S[best]) best = j;
    if (best != i) {
      std::swap(S[i], S[best]);
      for (int r = 0; r < 3; r++) {
        std::swap(U[r * 3 + i], U[r * 3 + best]);
        std::swap(V[r * 3 + i], V[r * 3 + best]);
      }
    }
  }
}

}  // namespace orc
