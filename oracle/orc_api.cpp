// ORACLE — TEST INFRASTRUCTURE ONLY.  C entry points (ctypes) over the CPU restatement in orc_*.h.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library.
// PARITY PIN: see oracle/README.md — the restatement is checked bit-for-bit against the reference's own
// translation units compiled in oracle/_ref (real SR/LO/LM sources + shim headers for ROS/PCL/OpenCV).
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>

#include "orc_cloud.h"
#include "orc_linalg.h"
#include "orc_lm.h"
#include "orc_lo.h"
#include "orc_sr.h"

using namespace orc;

namespace {

inline void to_cloud(const float* p, int n, Cloud& c) {
  c.resize(n);
  if (n) std::memcpy(c.data(), p, sizeof(P4) * (size_t)n);
}
inline int from_cloud(const Cloud& c, float* buf, int cap) {
  int n = (int)c.size();
  if (buf && cap >= n && n) std::memcpy(buf, c.data(), sizeof(P4) * (size_t)n);
  return n;
}
inline ImuTrans imu_from(const float* v) {
  ImuTrans t;
  if (v) {
    t.pitchStart = v[0]; t.yawStart = v[1]; t.rollStart = v[2];
    t.pitchLast = v[3]; t.yawLast = v[4]; t.rollLast = v[5];
    t.shiftX = v[6]; t.shiftY = v[7]; t.shiftZ = v[8];
    t.veloX = v[9]; t.veloY = v[10]; t.veloZ = v[11];
  }
  return t;
}
inline double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct SRH {
  SRParams prm;
  SRState st;
  SROut out;
  ImuState imu;
};

struct PipelineResult {  // mirrored by ctypes in tests/bench
  float odom[6];        // transformSum after this sweep (zeros on the init sweep)
  float mapped[6];      // transformAftMapped after the last mapping run
  float rel[6];         // sweep-relative transformation
  int odom_published, mapping_ran, odom_iters, map_iters;
  int n_full, n_sharp, n_less_sharp, n_flat, n_less_flat;
  int n_corner_stack, n_surf_stack, n_corner_map, n_surf_map;
  double t_extract, t_odom, t_map;  // seconds (steady_clock)
};

struct Pipeline {
  bool ros_hop = false;  // route the odometry pose through the quaternion message like the ROS nodes do
  SRH sr;
  LaserOdometry lo;
  LaserMapping lm;
  OdomOut oo;
  MapOut mo;
};

}  // namespace

extern "C" {

// ---------------------------------------------------------------- third-party restatements
int orc_voxel_grid(const float* in4, int m, float leaf, float* out4, int cap) {
  Cloud in, out;
  to_cloud(in4, m, in);
  voxel_grid(in, leaf, out);
  return from_cloud(out, out4, cap);
}

// idx/d2: nq*k entries, ascending (d2, idx); missing entries (cloud smaller than k) are -1 / inf.
int orc_knn(const float* cloud4, int n, const float* q4, int nq, int k, int brute, int* idx, float* d2) {
  Cloud c, q;
  to_cloud(cloud4, n, c);
  to_cloud(q4, nq, q);
  KnnIndex ki;
  ki.set(c, brute != 0);
  std::vector<Nbr> nb(k);
  for (int i = 0; i < nq; i++) {
    int f = ki.knn(q[i], k, nb.data());
    for (int j = 0; j < k; j++) {
      idx[i * k + j] = j < f ? nb[j].idx : -1;
      d2[i * k + j] = j < f ? nb[j].d2 : INFINITY;
    }
  }
  return 0;
}

void orc_gemm(const float* A, const float* B, float* C, int m, int k, int n) { gemm_f32_dacc(A, B, C, m, k, n); }
int orc_qr_solve(const float* A, const float* b, float* x, int m, int n) { return qr_solve(A, b, x, m, n) ? 0 : 1; }
void orc_jacobi_eigen(const float* A, float* W, float* V, int n) { jacobi_eigen(A, W, V, n); }
int orc_lu_inverse(const float* A, float* out, int n) { return lu_inverse(A, out, n) ? 0 : 1; }
void orc_svd3(const double* H, double* U, double* S, double* V) { svd3_jacobi(H, U, S, V); }
// state: 36 floats matP + 1 float degenerate flag
void orc_gn_solve(const float* AtA, const float* AtB, int iter, float thre, float* state37, float* X) {
  GNState st;
  std::memcpy(st.matP, state37, sizeof(st.matP));
  st.degenerate = state37[36] != 0.f;
  gn_solve_step(AtA, AtB, iter, thre, st, X);
  std::memcpy(state37, st.matP, sizeof(st.matP));
  state37[36] = st.degenerate ? 1.f : 0.f;
}

// ---------------------------------------------------------------- scanRegistration
void* orc_sr_create(int n_scans, int ring_mode, float ang_min, float ang_step) {
  SRH* h = new SRH;
  h->prm.n_scans = n_scans;
  h->prm.ring_mode = ring_mode;
  h->prm.ring_ang_min = ang_min;
  h->prm.ring_ang_step = ang_step;
  return h;
}
void orc_sr_destroy(void* h) { delete (SRH*)h; }
int orc_sr_extract(void* hv, const float* xyz, int n, int stride_floats) {
  SRH* h = (SRH*)hv;
  extract(h->prm, h->st, xyz, n, stride_floats, h->out);
  return 0;
}
// the IMU branch: one /imu/data message (SR:754-837); then extract with the sweep's stamp; imu_trans = the 12 floats of /imu_trans
void orc_sr_imu(void* hv, double stamp, const double* q4, const double* av3, const double* la3) { imu_handler(((SRH*)hv)->imu, stamp, q4, av3, la3); }
int orc_sr_extract_imu(void* hv, const float* xyz, int n, int stride_floats, double stamp, float* imu_trans12) {
  SRH* h = (SRH*)hv;
  extract(h->prm, h->st, xyz, n, stride_floats, h->out, &h->imu, stamp);
  if (imu_trans12) h->imu.trans12(imu_trans12);
  return 0;
}
// which: 0 full, 1 sharp, 2 less sharp, 3 flat, 4 less flat.  Returns the count (copies when cap suffices).
int orc_sr_cloud(void* hv, int which, float* buf, int cap) {
  SRH* h = (SRH*)hv;
  const Cloud* c[5] = {&h->out.full, &h->out.sharp, &h->out.lessSharp, &h->out.flat, &h->out.lessFlat};
  return from_cloud(*c[which], buf, cap);
}
// which: 0 scanStart, 1 scanEnd, 2 picked-after-mask, 3 label, 4 sortInd (3, 4 sized n_full)
int orc_sr_ints(void* hv, int which, int* buf, int cap) {
  SRH* h = (SRH*)hv;
  const int* src = nullptr;
  int n = 0;
  int nf = (int)h->out.full.size();
  switch (which) {
    case 0: src = h->out.scanStart.data(); n = (int)h->out.scanStart.size(); break;
    case 1: src = h->out.scanEnd.data(); n = (int)h->out.scanEnd.size(); break;
    case 2: src = h->out.pickedAfterMask.data(); n = (int)h->out.pickedAfterMask.size(); break;
    case 3: src = h->st.label.data(); n = nf; break;
    case 4: src = h->st.sortInd.data(); n = nf; break;
  }
  if (buf && cap >= n && n) std::memcpy(buf, src, sizeof(int) * (size_t)n);
  return n;
}
int orc_sr_curvature(void* hv, float* buf, int cap) {
  SRH* h = (SRH*)hv;
  int n = (int)h->out.full.size();
  if (buf && cap >= n && n) std::memcpy(buf, h->st.curvature.data(), sizeof(float) * (size_t)n);
  return n;
}

// ---------------------------------------------------------------- odometry, fine-grained
void orc_transform_to_start(const float* in4, int n, const float* T, float* out4) {
  for (int i = 0; i < n; i++) transform_to_start(T, ((const P4*)in4)[i], ((P4*)out4)[i]);
}
void orc_transform_to_end(const float* in4, int n, const float* T, const float* imu12, float* out4) {
  ImuTrans imu = imu_from(imu12);
  for (int i = 0; i < n; i++) transform_to_end(T, imu, ((const P4*)in4)[i], ((P4*)out4)[i]);
}
// One iteration body LO:586-974 without the solve.  c1..s3 are in/out (persist between iterations).
int orc_odom_iteration(const float* sharp, int n_sharp, const float* flat, int n_flat, const float* cornerLast, int n_cl,
                       const float* surfLast, int n_sl, const float* T, int iter, int brute, int* c1, int* c2, int* s1, int* s2,
                       int* s3, float* AtA36, float* AtB6, int* n_sel, float* rowsA, float* rowsB, int rows_cap) {
  Cloud cs, cf, ccl, csl;
  to_cloud(sharp, n_sharp, cs);
  to_cloud(flat, n_flat, cf);
  to_cloud(cornerLast, n_cl, ccl);
  to_cloud(surfLast, n_sl, csl);
  KnnIndex kc, ks;
  if (iter % 5 == 0) {
    kc.set(ccl, brute != 0);
    ks.set(csl, brute != 0);
  }
  OdomCorr corr;
  corr.c1.assign(c1, c1 + n_sharp); corr.c2.assign(c2, c2 + n_sharp);
  corr.s1.assign(s1, s1 + n_flat); corr.s2.assign(s2, s2 + n_flat); corr.s3.assign(s3, s3 + n_flat);
  NormalEq ne;
  odom_iteration(cs, cf, ccl, csl, kc, ks, T, iter, corr, ne);
  std::memcpy(c1, corr.c1.data(), sizeof(int) * n_sharp); std::memcpy(c2, corr.c2.data(), sizeof(int) * n_sharp);
  std::memcpy(s1, corr.s1.data(), sizeof(int) * n_flat); std::memcpy(s2, corr.s2.data(), sizeof(int) * n_flat);
  std::memcpy(s3, corr.s3.data(), sizeof(int) * n_flat);
  std::memcpy(AtA36, ne.AtA, sizeof(ne.AtA));
  std::memcpy(AtB6, ne.AtB, sizeof(ne.AtB));
  *n_sel = ne.n_sel;
  if (rowsA && rows_cap >= ne.n_sel && ne.n_sel >= 10) {
    std::memcpy(rowsA, ne.A.data(), sizeof(float) * 6 * ne.n_sel);
    std::memcpy(rowsB, ne.B.data(), sizeof(float) * ne.n_sel);
  }
  return 0;
}

// ---------------------------------------------------------------- mapping, fine-grained
void orc_associate_to_map(const float* in4, int n, const float* T, float* out4) {
  for (int i = 0; i < n; i++) associate_to_map(T, ((const P4*)in4)[i], ((P4*)out4)[i]);
}
void orc_associate_tobe_mapped(const float* in4, int n, const float* T, float* out4) {
  for (int i = 0; i < n; i++) associate_tobe_mapped(T, ((const P4*)in4)[i], ((P4*)out4)[i]);
}
void orc_transform_associate_to_map(const float* Tsum, const float* Tbef, const float* Taft, float* Tincre, float* Ttobe) {
  transform_associate_to_map(Tsum, Tbef, Taft, Tincre, Ttobe);
}
// One iteration body LM:754-967 without the solve.  corrC/corrS: 5 ints per stack point (-1 = rejected), may be NULL.
int orc_map_iteration(const float* cornerStack, int n_cs, const float* surfStack, int n_ss, const float* cornerMap, int n_cm,
                      const float* surfMap, int n_sm, const float* T, int brute, int* corrC, int* corrS, float* AtA36,
                      float* AtB6, int* n_sel) {
  Cloud a, b, c, d;
  to_cloud(cornerStack, n_cs, a);
  to_cloud(surfStack, n_ss, b);
  to_cloud(cornerMap, n_cm, c);
  to_cloud(surfMap, n_sm, d);
  KnnIndex kc, ks;
  kc.set(c, brute != 0);
  ks.set(d, brute != 0);
  MapCorr corr;
  NormalEq ne;
  map_iteration(a, b, c, d, kc, ks, T, &corr, ne);
  if (corrC && n_cs) std::memcpy(corrC, corr.corner.data(), sizeof(int) * 5 * n_cs);
  if (corrS && n_ss) std::memcpy(corrS, corr.surf.data(), sizeof(int) * 5 * n_ss);
  std::memcpy(AtA36, ne.AtA, sizeof(ne.AtA));
  std::memcpy(AtB6, ne.AtB, sizeof(ne.AtB));
  *n_sel = ne.n_sel;
  return 0;
}

// Same iteration, returning the 28 exact double sums {21 upper-triangle AtA, 6 AtB, n_sel} instead of the rounded
// matrices (what a rank contributes to the all-reduce when the map is sharded).
int orc_map_iteration_sums28(const float* cornerStack, int n_cs, const float* surfStack, int n_ss, const float* cornerMap, int n_cm,
                             const float* surfMap, int n_sm, const float* T, double* out28) {
  Cloud a, b, c, d;
  to_cloud(cornerStack, n_cs, a);
  to_cloud(surfStack, n_ss, b);
  to_cloud(cornerMap, n_cm, c);
  to_cloud(surfMap, n_sm, d);
  KnnIndex kc, ks;
  kc.set(c, false);
  ks.set(d, false);
  NormalEq ne;
  map_iteration(a, b, c, d, kc, ks, T, nullptr, ne, /*min_rows=*/0);
  int t = 0;
  for (int i = 0; i < 6; i++)
    for (int j = i; j < 6; j++) {
      double s = 0.0;
      for (int r = 0; r < ne.n_sel; r++) s += (double)ne.A[r * 6 + i] * (double)ne.A[r * 6 + j];
      out28[t++] = s;
    }
  for (int i = 0; i < 6; i++) {
    double s = 0.0;
    for (int r = 0; r < ne.n_sel; r++) s += (double)ne.A[r * 6 + i] * (double)ne.B[r];
    out28[21 + i] = s;
  }
  out28[27] = (double)ne.n_sel;
  return 0;
}

// ---------------------------------------------------------------- node level
void* orc_lo_create(int brute) {
  LaserOdometry* h = new LaserOdometry;
  h->brute = brute != 0;
  return h;
}
void orc_lo_destroy(void* h) { delete (LaserOdometry*)h; }
void orc_lo_control(void* h, int system_inited) { ((LaserOdometry*)h)->control(system_inited != 0); }

void* orc_lm_create(int brute) {
  LaserMapping* h = new LaserMapping;
  h->brute = brute != 0;
  return h;
}
void orc_lm_destroy(void* h) { delete (LaserMapping*)h; }
// One odometry message (LM:314-335) and, when `full_set`, one loop body LM:425-1139.  out: [0..6) aft, [6..12) bef,
// [12..18) tobe, then iterations, n_corner_map, n_surf_map, n_corner_stack, n_surf_stack, total corner, total surf.
int orc_lm_step(void* hv, const float* Tsum6, int full_set, const float* corner, int nc, const float* surf, int ns, const float* full, int nf,
                float* out25) {
  LaserMapping* lm = (LaserMapping*)hv;
  lm->odometry_msg(Tsum6);
  if (!full_set) return 0;
  Cloud c, s, f;
  to_cloud(corner, nc, c);
  to_cloud(surf, ns, s);
  to_cloud(full, nf, f);
  MapOut mo;
  lm->keepClouds = false;
  lm->process(c, s, f, mo);
  for (int i = 0; i < 6; i++) {
    out25[i] = mo.transformAftMapped[i];
    out25[6 + i] = mo.transformBefMapped[i];
    out25[12 + i] = mo.transformTobeMapped[i];
  }
  size_t a = 0, b = 0;
  for (auto& cl : lm->cornerArr) a += cl.size();
  for (auto& cl : lm->surfArr) b += cl.size();
  out25[18] = (float)mo.iterations; out25[19] = (float)mo.nCornerFromMap; out25[20] = (float)mo.nSurfFromMap;
  out25[21] = (float)mo.nCornerStack; out25[22] = (float)mo.nSurfStack; out25[23] = (float)a; out25[24] = (float)b;
  return 0;
}

// ---------------------------------------------------------------- transformMaintenance (N2)
void* orc_tm_create() { return new TransformMaintenance; }
void orc_tm_destroy(void* h) { delete (TransformMaintenance*)h; }
void orc_tm_odometry(void* h, const float* Tsum6, double stamp, float* out6, double* track4) { ((TransformMaintenance*)h)->odometry(Tsum6, stamp, out6, track4); }
void orc_tm_aft_mapped(void* h, const float* aft6, const float* bef6) { ((TransformMaintenance*)h)->aft_mapped(aft6, bef6); }

// ---------------------------------------------------------------- full pipeline SR -> LO -> LM (one process, no ROS)
void* orc_pipeline_create(int n_scans, int ring_mode, float ang_min, float ang_step, int brute, int keep_clouds) {
  Pipeline* p = new Pipeline;
  p->sr.prm.n_scans = n_scans;
  p->sr.prm.ring_mode = ring_mode;
  p->sr.prm.ring_ang_min = ang_min;
  p->sr.prm.ring_ang_step = ang_step;
  p->lo.brute = brute != 0;
  p->lm.brute = brute != 0;
  p->lm.keepClouds = keep_clouds != 0;
  return p;
}
void orc_pipeline_destroy(void* h) { delete (Pipeline*)h; }

// The odometry node alone (one main-loop body LO:502-1147 per call) with the twelve floats of /imu_trans as input: pins the
// IMU formulas of laserOdometry (LO:201-225, 385-409, 566-568, 1053-1064) against the reference's own laserOdometry.cpp.
// out18 = transformSum[6], transformation[6], {odometry published, clouds published, full-res published, iterations, 0, 0}.
struct LOH {
  LaserOdometry lo;
  OdomOut oo;
};
void* orc_lonode_create() { return new LOH(); }
void orc_lonode_destroy(void* h) { delete (LOH*)h; }
void orc_lonode_control(void* h, int inited) { ((LOH*)h)->lo.control(inited != 0); }
int orc_lonode_step(void* hv, const float* sharp, int ns, const float* less_sharp, int nls, const float* flat, int nf, const float* less_flat,
                int nlf, const float* full, int nfull, const float* imu12, float* out18) {
  LOH* h = (LOH*)hv;
  Cloud a, b, c, d, e;
  to_cloud(sharp, ns, a);
  to_cloud(less_sharp, nls, b);
  to_cloud(flat, nf, c);
  to_cloud(less_flat, nlf, d);
  to_cloud(full, nfull, e);
  h->lo.process(a, b, c, d, e, imu_from(imu12), h->oo);
  for (int i = 0; i < 6; i++) {
    out18[i] = h->oo.transformSum[i];
    out18[6 + i] = h->oo.transformation[i];
  }
  out18[12] = h->oo.odomPublished ? 1.f : 0.f;
  out18[13] = h->oo.cloudsPublished ? 1.f : 0.f;
  out18[14] = h->oo.fullResPublished ? 1.f : 0.f;
  out18[15] = (float)h->oo.iterations;
  out18[16] = out18[17] = 0.f;
  return 0;
}
// which: 0 /laser_cloud_corner_last, 1 /laser_cloud_surf_last, 2 /velodyne_cloud_3 (valid when the flags of the step say so)
int orc_lonode_cloud(void* hv, int which, float* buf, int cap) {
  LOH* h = (LOH*)hv;
  return from_cloud(which == 0 ? h->oo.cornerLast : (which == 1 ? h->oo.surfLast : h->oo.fullRes), buf, cap);
}
void orc_pipeline_set_ros_hop(void* h, int on) { ((Pipeline*)h)->ros_hop = on != 0; }
void orc_odometry_ros_hop(const float* in6, float* out6) { odometry_ros_hop(in6, out6); }
// IMControl{systemInited=false} (IN:281-284): odometry re-initialises on the next sweep, mapping when it sees zero odometry.
void orc_pipeline_reset(void* h) { ((Pipeline*)h)->lo.control(false); }

int orc_pipeline_process(void* hv, const float* xyz, int n, int stride_floats, PipelineResult* r) {
  Pipeline* p = (Pipeline*)hv;
  std::memset(r, 0, sizeof(*r));
  double t0 = now_s();
  extract(p->sr.prm, p->sr.st, xyz, n, stride_floats, p->sr.out);
  double t1 = now_s();
  const SROut& f = p->sr.out;
  ImuTrans imu;
  p->lo.process(f.sharp, f.lessSharp, f.flat, f.lessFlat, f.full, imu, p->oo);
  double t2 = now_s();
  r->odom_published = p->oo.odomPublished;
  if (p->oo.odomPublished) {
    float ts[6];
    if (p->ros_hop) odometry_ros_hop(p->oo.transformSum, ts); else std::memcpy(ts, p->oo.transformSum, sizeof(ts));
    p->lm.odometry_msg(ts);
  }
  if (p->oo.odomPublished && p->oo.fullResPublished) {
    p->lm.process(p->oo.cornerLast, p->oo.surfLast, p->oo.fullRes, p->mo);
    r->mapping_ran = 1;
    r->map_iters = p->mo.iterations;
    r->n_corner_stack = p->mo.nCornerStack; r->n_surf_stack = p->mo.nSurfStack;
    r->n_corner_map = p->mo.nCornerFromMap; r->n_surf_map = p->mo.nSurfFromMap;
  }
  double t3 = now_s();
  for (int i = 0; i < 6; i++) {
    r->odom[i] = p->oo.transformSum[i];
    r->rel[i] = p->oo.transformation[i];
    r->mapped[i] = p->lm.Taft[i];
  }
  r->odom_iters = p->oo.iterations;
  r->n_full = (int)f.full.size(); r->n_sharp = (int)f.sharp.size(); r->n_less_sharp = (int)f.lessSharp.size();
  r->n_flat = (int)f.flat.size(); r->n_less_flat = (int)f.lessFlat.size();
  r->t_extract = t1 - t0; r->t_odom = t2 - t1; r->t_map = t3 - t2;
  return 0;
}
// transformBefMapped as /aft_mapped_to_init carries it in the twist fields (LM:1125-1130): transformMaintenance needs it
int orc_pipeline_bef_mapped(void* hv, float* out6) {
  Pipeline* p = (Pipeline*)hv;
  for (int i = 0; i < 6; i++) out6[i] = p->lm.Tbef[i];
  return 0;
}
// which: 0..4 SR clouds (as orc_sr_cloud), 5 cornerLast, 6 surfLast, 7 fullRes (odometry outputs, valid when published),
// 8 surround, 9 fullResRegistered (mapping outputs)
int orc_pipeline_cloud(void* hv, int which, float* buf, int cap) {
  Pipeline* p = (Pipeline*)hv;
  switch (which) {
    case 0: return from_cloud(p->sr.out.full, buf, cap);
    case 1: return from_cloud(p->sr.out.sharp, buf, cap);
    case 2: return from_cloud(p->sr.out.lessSharp, buf, cap);
    case 3: return from_cloud(p->sr.out.flat, buf, cap);
    case 4: return from_cloud(p->sr.out.lessFlat, buf, cap);
    case 5: return from_cloud(p->oo.cornerLast, buf, cap);
    case 6: return from_cloud(p->oo.surfLast, buf, cap);
    case 7: return from_cloud(p->oo.fullRes, buf, cap);
    case 8: return from_cloud(p->mo.surround, buf, cap);
    case 9: return from_cloud(p->mo.fullResRegistered, buf, cap);
  }
  return -1;
}
// Total number of points held in the mapping cube grid (corner, surf).
void orc_pipeline_map_size(void* hv, int* n_corner, int* n_surf) {
  Pipeline* p = (Pipeline*)hv;
  size_t a = 0, b = 0;
  for (auto& c : p->lm.cornerArr) a += c.size();
  for (auto& c : p->lm.surfArr) b += c.size();
  *n_corner = (int)a;
  *n_surf = (int)b;
}

}  // extern "C"

// ---------------------------------------------------------------- the reference's process layout: three single-threaded
// stages (scanRegistration | laserOdometry | laserMapping are separate ROS processes) connected by queues.
// Same results as orc_pipeline_process (asserted in tests); used by bench.py --impl reference to give the CPU path
// all the host threads its structure can use for ONE sequence.
namespace {
template <typename T>
struct Chan {
  std::mutex m;
  std::condition_variable cv;
  std::deque<T> q;
  bool closed = false;
  void push(T&& v) {
    { std::lock_guard<std::mutex> l(m); q.push_back(std::move(v)); }
    cv.notify_one();
  }
  void close() {
    { std::lock_guard<std::mutex> l(m); closed = true; }
    cv.notify_all();
  }
  bool pop(T& out) {
    std::unique_lock<std::mutex> l(m);
    cv.wait(l, [&] { return !q.empty() || closed; });
    if (q.empty()) return false;
    out = std::move(q.front());
    q.pop_front();
    return true;
  }
};
struct Msg1 { int k; SROut f; };
struct Msg2 { int k; bool odom; bool full; float Tsum[6]; Cloud corner, surf, fullres; };
}  // namespace

extern "C" int orc_pipeline_run_threaded(void* hv, const float* xyz_all, const long long* offsets, int n_sweeps, PipelineResult* res) {
  Pipeline* p = (Pipeline*)hv;
  for (int k = 0; k < n_sweeps; k++) std::memset(&res[k], 0, sizeof(PipelineResult));
  Chan<Msg1> c1;
  Chan<Msg2> c2;
  std::thread tA([&] {
    for (int k = 0; k < n_sweeps; k++) {
      double t0 = now_s();
      Msg1 m;
      m.k = k;
      extract(p->sr.prm, p->sr.st, xyz_all + 3 * offsets[k], (int)(offsets[k + 1] - offsets[k]), 3, m.f);
      res[k].t_extract = now_s() - t0;
      res[k].n_full = (int)m.f.full.size(); res[k].n_sharp = (int)m.f.sharp.size(); res[k].n_less_sharp = (int)m.f.lessSharp.size();
      res[k].n_flat = (int)m.f.flat.size(); res[k].n_less_flat = (int)m.f.lessFlat.size();
      c1.push(std::move(m));
    }
    c1.close();
  });
  std::thread tB([&] {
    Msg1 m;
    ImuTrans imu;
    OdomOut oo;
    while (c1.pop(m)) {
      double t0 = now_s();
      p->lo.process(m.f.sharp, m.f.lessSharp, m.f.flat, m.f.lessFlat, m.f.full, imu, oo);
      PipelineResult& r = res[m.k];
      r.t_odom = now_s() - t0;
      r.odom_published = oo.odomPublished;
      r.odom_iters = oo.iterations;
      for (int i = 0; i < 6; i++) { r.odom[i] = oo.transformSum[i]; r.rel[i] = oo.transformation[i]; }
      Msg2 o;
      o.k = m.k;
      o.odom = oo.odomPublished;
      o.full = oo.odomPublished && oo.fullResPublished;
      for (int i = 0; i < 6; i++) o.Tsum[i] = oo.transformSum[i];
      if (o.full) { o.corner = oo.cornerLast; o.surf = oo.surfLast; o.fullres = oo.fullRes; }
      c2.push(std::move(o));
    }
    c2.close();
  });
  std::thread tC([&] {
    Msg2 o;
    MapOut mo;
    while (c2.pop(o)) {
      double t0 = now_s();
      PipelineResult& r = res[o.k];
      if (o.odom) p->lm.odometry_msg(o.Tsum);
      if (o.full) {
        p->lm.process(o.corner, o.surf, o.fullres, mo);
        r.mapping_ran = 1;
        r.map_iters = mo.iterations;
        r.n_corner_stack = mo.nCornerStack; r.n_surf_stack = mo.nSurfStack;
        r.n_corner_map = mo.nCornerFromMap; r.n_surf_map = mo.nSurfFromMap;
      }
      for (int i = 0; i < 6; i++) r.mapped[i] = p->lm.Taft[i];
      r.t_map = now_s() - t0;
    }
  });
  tA.join();
  tB.join();
  tC.join();
  return 0;
}
