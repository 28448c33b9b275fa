// N4: track calibration -- weighted rigid alignment of a SLAM track to its GPS (ENU) track, the
// re-weighting loop around it and the pairwise smoothing that follows.  Replaces the bodies of
//   trackCalibration            src/gpsCalibration/src/gps_calibration/track_calibration.cc   (TC)
//   WeightCoeCal                src/gpsCalibration/src/gps_calibration/weight_calculation.cc  (WC)
//   longDisTrackPro (the loop)  src/gpsCalibration/src/long_distance_track_process/long_distance_track_process.cpp (LD:57-83)
// north_star keeps the trajectory alignment in host C++: everything here is fp64 host arithmetic in the reference's
// operation order (sequential sums, no FMA contraction), except the O(N^2) smoothing loop (TC:648-674), which runs as
// one kernel -- a thread per track point walks the whole track in the reference's order, so every point gets the very
// sum the serial loop gives -- or, on request, as its O(N) closed form on the host.
// Third-party piece: Eigen's JacobiSVD of the 3x3 cross-covariance (TC:506) is restated as a two-sided Jacobi SVD.
#include <cfloat>
#include <cmath>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/loamgpu.h"
#include "lg_common.cuh"

namespace {

// ------------------------------------------------------------------------------------------------ 3x3 helpers
struct M3 {
  double a[3][3];
};
inline M3 m3_identity() {
  M3 r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) r.a[i][j] = i == j ? 1.0 : 0.0;
  return r;
}
inline M3 m3_mul(const M3& x, const M3& y) {
  M3 r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double s = x.a[i][0] * y.a[0][j];
      s += x.a[i][1] * y.a[1][j];
      s += x.a[i][2] * y.a[2][j];
      r.a[i][j] = s;
    }
  return r;
}
inline M3 m3_t(const M3& x) {
  M3 r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) r.a[i][j] = x.a[j][i];
  return r;
}
inline double m3_det(const M3& m) {
  return m.a[0][0] * (m.a[1][1] * m.a[2][2] - m.a[1][2] * m.a[2][1]) - m.a[0][1] * (m.a[1][0] * m.a[2][2] - m.a[1][2] * m.a[2][0]) +
         m.a[0][2] * (m.a[1][0] * m.a[2][1] - m.a[1][1] * m.a[2][0]);
}

// Two-sided Jacobi SVD of a real 3x3 matrix, H = U diag(S) V^T, S descending and non-negative.  Every (p, q) block is
// first made symmetric by a left rotation and then diagonalised by a symmetric Jacobi rotation; rows / columns that
// are exactly zero are never touched, so a planar problem (third row and column zero, TC:52-57) keeps U(:,2) = V(:,2) = e3.
void svd3(const M3& H, M3& U, double S[3], M3& V) {
  M3 W = H;
  U = m3_identity();
  V = m3_identity();
  for (int sweep = 0; sweep < 60; sweep++) {
    bool rotated = false;
    for (int q = 1; q < 3; q++)
      for (int p = 0; p < q; p++) {
        double dmax = fmax(fabs(W.a[0][0]), fmax(fabs(W.a[1][1]), fabs(W.a[2][2])));
        double thr = fmax(DBL_MIN, DBL_EPSILON * dmax);
        if (!(fabs(W.a[p][q]) > thr || fabs(W.a[q][p]) > thr)) continue;
        rotated = true;
        const double a = W.a[p][p], b = W.a[p][q], c = W.a[q][p], d = W.a[q][q];
        // left rotation G = [[c1, s1], [-s1, c1]] with G * block symmetric
        double c1 = 1.0, s1 = 0.0;
        const double t = a + d, dd = c - b;
        if (fabs(dd) >= DBL_MIN) {
          const double u = t / dd, tmp = sqrt(1.0 + u * u);
          s1 = 1.0 / tmp;
          c1 = u / tmp;
        }
        const double x = c1 * a + s1 * c, y = c1 * b + s1 * d, z = -s1 * b + c1 * d;
        // Jacobi rotation J = [[cj, sj], [-sj, cj]] with J^T [[x, y], [y, z]] J diagonal
        double cj = 1.0, sj = 0.0;
        if (fabs(y) >= DBL_MIN) {
          const double tau = (z - x) / (2.0 * y);
          const double tt = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
          cj = 1.0 / sqrt(1.0 + tt * tt);
          sj = tt * cj;
        }
        M3 G = m3_identity(), J = m3_identity();
        G.a[p][p] = c1; G.a[p][q] = s1; G.a[q][p] = -s1; G.a[q][q] = c1;
        J.a[p][p] = cj; J.a[p][q] = sj; J.a[q][p] = -sj; J.a[q][q] = cj;
        const M3 L = m3_mul(m3_t(G), J);  // block = L diag L^T-ish: W' = L^T W J
        W = m3_mul(m3_mul(m3_t(L), W), J);
        W.a[p][q] = 0.0;
        W.a[q][p] = 0.0;
        U = m3_mul(U, L);
        V = m3_mul(V, J);
      }
    if (!rotated) break;
  }
  for (int i = 0; i < 3; i++) {
    S[i] = fabs(W.a[i][i]);
    if (W.a[i][i] < 0.0)
      for (int r = 0; r < 3; r++) U.a[r][i] = -U.a[r][i];
  }
  for (int i = 0; i < 2; i++) {  // descending, stable
    int best = i;
    for (int j = i + 1; j < 3; j++)
      if (S[j] > S[best]) best = j;
    if (best != i) {
      double ts = S[i]; S[i] = S[best]; S[best] = ts;
      for (int r = 0; r < 3; r++) {
        double tu = U.a[r][i]; U.a[r][i] = U.a[r][best]; U.a[r][best] = tu;
        double tv = V.a[r][i]; V.a[r][i] = V.a[r][best]; V.a[r][best] = tv;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ TC
struct Track {  // homogeneous N x 4 rows {x, y, 1, 1} like SLAMCoord / ENUCoord (TC:52-68)
  std::vector<double> v;
  int n = 0;
  void init(int n_) { n = n_; v.assign((size_t)n * 4, 1.0); }
  double& at(int i, int j) { return v[(size_t)i * 4 + j]; }
  double at(int i, int j) const { return v[(size_t)i * 4 + j]; }
};

// BFTWithWeight TC:366-545: weighted centroids, weighted cross-covariance, SVD, reflection fix, homogeneous transform
void best_fit_weighted(const Track& A, const Track& B, const double* w, double T[16]) {
  const int n = A.n;
  double sa[3] = {0, 0, 0}, sb[3] = {0, 0, 0}, sw = 0.0;
  for (int i = 0; i < n; i++) {  // TC:427-439 (sums of the pre-weighted copies A1 / B1, TC:417-425)
    for (int j = 0; j < 3; j++) {
      sa[j] += A.at(i, j) * w[i];
      sb[j] += B.at(i, j) * w[i];
    }
    sw += w[i];
  }
  for (int j = 0; j < 3; j++) {  // TC:450-456
    sa[j] = sa[j] / sw;
    sb[j] = sb[j] / sw;
  }
  M3 H;
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) H.a[r][c] = 0.0;
  for (int i = 0; i < n; i++) {  // TC:489-503: H = AA^T * BB, rows weighted after centring
    double aa[3], bb[3];
    for (int j = 0; j < 3; j++) {
      aa[j] = (A.at(i, j) - sa[j]) * w[i];
      bb[j] = (B.at(i, j) - sb[j]) * w[i];
    }
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) H.a[r][c] += aa[r] * bb[c];
  }
  M3 U, V;
  double S[3];
  svd3(H, U, S, V);                      // TC:506-509
  M3 R = m3_mul(V, m3_t(U));             // TC:511
  if (m3_det(R) < 0) {                   // TC:514-521
    for (int i = 0; i < 3; i++) V.a[i][2] = -1 * V.a[i][2];
    R = m3_mul(V, m3_t(U));
  }
  for (int i = 0; i < 16; i++) T[i] = (i % 5 == 0) ? 1.0 : 0.0;
  for (int i = 0; i < 3; i++) {          // TC:524-539
    double ra = R.a[i][0] * sa[0];
    ra += R.a[i][1] * sa[1];
    ra += R.a[i][2] * sa[2];
    for (int j = 0; j < 3; j++) T[i * 4 + j] = R.a[i][j];
    T[i * 4 + 3] = sb[i] - ra;
  }
}

// icp TC:98-201 + coordRotated TC:583-618
void icp_rotate(const Track& slam, const Track& enu, const double* w, double T[16], double* rotated_xy) {
  const int n = slam.n;
  Track src = slam;  // TC:123-134: columns 0..2 copied, column 3 stays 1
  std::vector<double> dist(n);
  double prev = 0.0;
  for (int it = 0; it < 2; it++) {  // TC:147 maxIterations = 2
    for (int i = 0; i < n; i++) {   // nearestNeighbor TC:557-579: index i pairs with index i
      double dx = src.at(i, 0) - enu.at(i, 0), dy = src.at(i, 1) - enu.at(i, 1);
      dist[i] = sqrt(dx * dx + dy * dy);
    }
    double Ti[16];
    best_fit_weighted(src, enu, w, Ti);
    for (int i = 0; i < n; i++) {  // TC:165 src = src * T^T
      double row[4];
      for (int j = 0; j < 4; j++) {
        double s = src.at(i, 0) * Ti[j * 4 + 0];
        s += src.at(i, 1) * Ti[j * 4 + 1];
        s += src.at(i, 2) * Ti[j * 4 + 2];
        s += src.at(i, 3) * Ti[j * 4 + 3];
        row[j] = s;
      }
      for (int j = 0; j < 4; j++) src.at(i, j) = row[j];
    }
    double mean = 0.0;
    for (int i = 0; i < n; i++) mean += dist[i];
    mean = mean / n;
    if (fabs(prev - mean) < 0.003) break;  // TC:176
    prev = mean;
  }
  best_fit_weighted(slam, src, w, T);  // TC:187-189
  for (int i = 0; i < n; i++)          // TC:615: SLAMRotatedCoord = SLAM * R^T + t
    for (int j = 0; j < 2; j++) {
      double s = slam.at(i, 0) * T[j * 4 + 0];
      s += slam.at(i, 1) * T[j * 4 + 1];
      s += slam.at(i, 2) * T[j * 4 + 2];
      rotated_xy[2 * i + j] = s + T[j * 4 + 3];
    }
}

void load_tracks(const double* slam_xyzt, const double* enu_xyzt, int n, Track& slam, Track& enu) {  // dataInitial TC:52-68
  slam.init(n);
  enu.init(n);
  for (int i = 0; i < n; i++) {
    slam.at(i, 0) = slam_xyzt[4 * i] - slam_xyzt[0];
    slam.at(i, 1) = slam_xyzt[4 * i + 1] - slam_xyzt[1];
    enu.at(i, 0) = enu_xyzt[4 * i] - enu_xyzt[0];
    enu.at(i, 1) = enu_xyzt[4 * i + 1] - enu_xyzt[1];
  }
}

// ------------------------------------------------------------------------------------------------ TC:631-689 on the device
// One thread per track point iNum; the track (rotated SLAM xy, ENU xy) is staged through shared memory in tiles and
// walked in index order, so thread iNum performs exactly the serial loop's additions TC:654-663.
constexpr int TS_THREADS = 128;
__global__ void __launch_bounds__(TS_THREADS) track_smooth_kernel(const double2* __restrict__ rot, const double2* __restrict__ enu, int n,
                                                                  double2* __restrict__ out) {
  __shared__ double2 s_rot[TS_THREADS], s_enu[TS_THREADS];
  const int i = blockIdx.x * TS_THREADS + threadIdx.x;
  const double2 me = i < n ? rot[i] : make_double2(0.0, 0.0);
  double ax = 0.0, ay = 0.0;
  for (int base = 0; base < n; base += TS_THREADS) {
    const int j = base + threadIdx.x;
    if (j < n) {
      s_rot[threadIdx.x] = rot[j];
      s_enu[threadIdx.x] = enu[j];
    }
    __syncthreads();
    const int m = min(TS_THREADS, n - base);
    for (int k = 0; k < m; k++) {
      const double dx = s_rot[k].x - me.x, dy = s_rot[k].y - me.y;  // TC:657-658
      ax += s_enu[k].x - dx;                                         // TC:661-662
      ay += s_enu[k].y - dy;
    }
    __syncthreads();
  }
  if (i < n) {
    ax /= n;  // TC:666-667
    ay /= n;
    out[i] = make_double2((ax + me.x) / 2.0, (ay + me.y) / 2.0);  // TC:670-671
  }
}

int smooth_device(const double* rot_xy, const double* enu_rel_xy, int n, int device, double* out_xy) {
  if (cudaSetDevice(device) != cudaSuccess) {
    lg_set_error("loam_track_smooth: no CUDA device (mode 0 runs on the GPU; mode 1 is the host closed form)", __FILE__, __LINE__);
    return LOAM_ECUDA;
  }
  double2 *d_rot = nullptr, *d_enu = nullptr, *d_out = nullptr;
  const size_t bytes = (size_t)n * sizeof(double2);
  int rc = LOAM_OK;
  cudaError_t e = cudaMalloc((void**)&d_rot, bytes);
  if (e == cudaSuccess) e = cudaMalloc((void**)&d_enu, bytes);
  if (e == cudaSuccess) e = cudaMalloc((void**)&d_out, bytes);
  if (e == cudaSuccess) e = cudaMemcpy(d_rot, rot_xy, bytes, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_enu, enu_rel_xy, bytes, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    track_smooth_kernel<<<lg_div_up(n, TS_THREADS), TS_THREADS>>>(d_rot, d_enu, n, d_out);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(out_xy, d_out, bytes, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) {
    lg_set_error(cudaGetErrorString(e), __FILE__, __LINE__);
    rc = LOAM_ECUDA;
  }
  cudaFree(d_rot);
  cudaFree(d_enu);
  cudaFree(d_out);
  return rc;
}

// O(N) closed form of the same loop: sum_i (E_i - (S_i - S_k)) / N = mean(E) - mean(S) + S_k.  Not bit-identical to the
// serial sum (different rounding), agrees to ~1e-12 relative.
void smooth_closed_form(const double* rot_xy, const double* enu_rel_xy, int n, double* out_xy) {
  double se[2] = {0, 0}, ss[2] = {0, 0};
  for (int i = 0; i < n; i++)
    for (int j = 0; j < 2; j++) {
      se[j] += enu_rel_xy[2 * i + j];
      ss[j] += rot_xy[2 * i + j];
    }
  for (int k = 0; k < n; k++)
    for (int j = 0; j < 2; j++) {
      double avg = (se[j] - ss[j]) / n + rot_xy[2 * k + j];
      out_xy[2 * k + j] = (avg + rot_xy[2 * k + j]) / 2.0;
    }
}

}  // namespace

extern "C" {

int loam_track_svd3(const double* h9, double* u9, double* s3, double* v9) {
  if (!h9 || !u9 || !s3 || !v9) return LOAM_EINVAL;
  M3 H, U, V;
  memcpy(H.a, h9, sizeof(H.a));
  svd3(H, U, s3, V);
  memcpy(u9, U.a, sizeof(U.a));
  memcpy(v9, V.a, sizeof(V.a));
  return LOAM_OK;
}

int loam_track_speed_weights(const double* slam_xyzt, int n, double* w) {
  if (!slam_xyzt || !w || n < 1) return LOAM_EINVAL;
  for (int is = 0; is < n; is++) {
    if (is == 0) {
      w[is] = 1.0;
      continue;
    }
    const int nx = is + 1 < n ? is + 1 : n - 1;  // quirk fence: WC:18-19 reads element n for the last point (out of bounds)
    double dx = slam_xyzt[4 * nx] - slam_xyzt[4 * is], dy = slam_xyzt[4 * nx + 1] - slam_xyzt[4 * is + 1];
    double dis = sqrt(dx * dx + dy * dy);
    w[is] = fmin(dis / 2.2, 1.0);  // SPEED, weight_calculation.h:6
  }
  return LOAM_OK;
}

int loam_track_residual_weights(const double* slam_xyzt, const double* enu_xyzt, const double* cal_xyzt, int n, double* w) {
  if (!enu_xyzt || !cal_xyzt) return LOAM_EINVAL;
  int rc = loam_track_speed_weights(slam_xyzt, n, w);  // WC:35-48
  if (rc) return rc;
  for (int is = 0; is < n; is++) {  // WC:68-75 (maxDis / minDis, WC:50-61, are computed and never used)
    double dx = enu_xyzt[4 * is] - cal_xyzt[4 * is], dy = enu_xyzt[4 * is + 1] - cal_xyzt[4 * is + 1];
    double dis = sqrt(dx * dx + dy * dy);
    w[is] = w[is] * 1.0 / fmax(0.01, dis);  // DELTA, weight_calculation.h:7
  }
  return LOAM_OK;
}

int loam_track_icp(const double* slam_xyzt, const double* enu_xyzt, const double* w, int n, double* T16, double* rotated_xy) {
  if (!slam_xyzt || !enu_xyzt || !w || !T16 || !rotated_xy || n < 1) return LOAM_EINVAL;
  Track slam, enu;
  load_tracks(slam_xyzt, enu_xyzt, n, slam, enu);
  icp_rotate(slam, enu, w, T16, rotated_xy);
  return LOAM_OK;
}

int loam_track_smooth(const double* rotated_xy, const double* enu_xyzt, int n, int mode, int device, double* cal_xyzt) {
  if (!rotated_xy || !enu_xyzt || !cal_xyzt || n < 1 || (mode != 0 && mode != 1)) return LOAM_EINVAL;
  std::vector<double> enu_rel((size_t)n * 2), out((size_t)n * 2);
  for (int i = 0; i < n; i++) {
    enu_rel[2 * i] = enu_xyzt[4 * i] - enu_xyzt[0];
    enu_rel[2 * i + 1] = enu_xyzt[4 * i + 1] - enu_xyzt[1];
  }
  if (mode == 0) {
    int rc = smooth_device(rotated_xy, enu_rel.data(), n, device, out.data());
    if (rc) return rc;
  } else {
    smooth_closed_form(rotated_xy, enu_rel.data(), n, out.data());
  }
  for (int i = 0; i < n; i++) {  // TC:676-686
    cal_xyzt[4 * i] = out[2 * i] + enu_xyzt[0];
    cal_xyzt[4 * i + 1] = out[2 * i + 1] + enu_xyzt[1];
    cal_xyzt[4 * i + 2] = enu_xyzt[4 * i + 2];
    cal_xyzt[4 * i + 3] = enu_xyzt[4 * i + 3];
  }
  return LOAM_OK;
}

int loam_track_calibrate(const double* slam_xyzt, const double* enu_xyzt, const double* w, int n, int mode, int device, double* cal_xyzt,
                         double* T16) {
  if (!cal_xyzt || n < 1) return LOAM_EINVAL;
  std::vector<double> rot((size_t)n * 2);
  double T[16];
  int rc = loam_track_icp(slam_xyzt, enu_xyzt, w, n, T, rot.data());
  if (rc) return rc;
  if (T16) memcpy(T16, T, sizeof(T));
  return loam_track_smooth(rot.data(), enu_xyzt, n, mode, device, cal_xyzt);
}

int loam_track_calibrate_long(const double* slam_xyzt, const double* enu_xyzt, int n, int iterations, int mode, int device, double* w_out,
                              double* cal_xyzt) {
  if (!slam_xyzt || !enu_xyzt || !w_out || !cal_xyzt || n < 1 || iterations < 0) return LOAM_EINVAL;
  int rc = loam_track_speed_weights(slam_xyzt, n, w_out);  // LD:61-63
  if (rc) return rc;
  rc = loam_track_calibrate(slam_xyzt, enu_xyzt, w_out, n, mode, device, cal_xyzt, nullptr);  // LD:67-72
  if (rc) return rc;
  std::vector<double> prev((size_t)n * 4);
  for (int i = 1; i <= iterations; i++) {  // LD:74-84 (MAXITERATOR = 5)
    rc = loam_track_residual_weights(slam_xyzt, enu_xyzt, cal_xyzt, n, w_out);
    if (rc) return rc;
    memcpy(prev.data(), cal_xyzt, prev.size() * sizeof(double));
    rc = loam_track_calibrate(prev.data(), enu_xyzt, w_out, n, mode, device, cal_xyzt, nullptr);  // the calibrated track is the new "SLAM" track
    if (rc) return rc;
  }
  return LOAM_OK;
}

}  // extern "C"
