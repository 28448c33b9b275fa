// ORACLE — TEST INFRASTRUCTURE ONLY.  Compiles the reference's track_calibration.cc and weight_calculation.cc unmodified
// (read in place from /root/reference) against the minimal Eigen shim (shim/eigenshim.h).  `private` is opened so the
// harness can read SLAMRotatedCoord; nothing in the sources changes.
// WC:18-19 / WC:41-42 read SLAMTrackTmp[n] for the last point (out of bounds): the harness keeps a copy of the last
// element behind the vector's end (capacity n + 1), which is the clamp the product and the oracle state as their fence.
#include "ref_common.h"
#include <algorithm>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <sstream>
#include "common.h"
#include "shim/eigenshim.h"
#define private public
#include "track_calibration.h"
#undef private
#include "track_calibration.cc"
#include "weight_calculation.cc"

static std::vector<COORDXYZT> make_track(const double* xyzt, int n) {
  std::vector<COORDXYZT> v;
  v.reserve((size_t)n + 1);
  for (int i = 0; i < n; i++) v.push_back(COORDXYZT{xyzt[4 * i], xyzt[4 * i + 1], xyzt[4 * i + 2], xyzt[4 * i + 3]});
  if (n > 0) v.data()[n] = v.data()[n - 1];  // behind the end, inside the capacity
  return v;
}
static void put_track(const std::vector<COORDXYZT>& v, double* out) {
  for (size_t i = 0; i < v.size(); i++) {
    out[4 * i] = v[i].x; out[4 * i + 1] = v[i].y; out[4 * i + 2] = v[i].z; out[4 * i + 3] = v[i].t;
  }
}

REF_API int ref_wc_speed(const double* slam, int n, double* w) {
  std::vector<COORDXYZT> s = make_track(slam, n);
  std::vector<double> wc;
  WeightCoeCal().ICPWeightCoeCal(s, wc);
  for (int i = 0; i < n; i++) w[i] = wc[i];
  return (int)wc.size();
}
REF_API int ref_wc_residual(const double* slam, const double* enu, const double* cal, int n, double* w) {
  std::vector<COORDXYZT> s = make_track(slam, n), e = make_track(enu, n), c = make_track(cal, n);
  std::vector<double> wc;
  WeightCoeCal().ICPWeightCoeCal(s, wc, e, c);
  for (int i = 0; i < n; i++) w[i] = wc[i];
  return (int)wc.size();
}
// one trackCalibration object: constructor + doICP + doCalibration (SD:241-243); rotated_xy = SLAMRotatedCoord(:, 0..1)
REF_API int ref_tc_calibrate(const double* slam, const double* enu, const double* w, int n, double* cal, double* rotated_xy) {
  std::vector<COORDXYZT> s = make_track(slam, n), e = make_track(enu, n);
  std::vector<double> wc(w, w + n);
  trackCalibration tc(s, e, wc);
  tc.doICP();
  if (rotated_xy)
    for (int i = 0; i < n; i++) {
      rotated_xy[2 * i] = tc.SLAMRotatedCoord(i, 0);
      rotated_xy[2 * i + 1] = tc.SLAMRotatedCoord(i, 1);
    }
  std::vector<COORDXYZT> out;
  tc.doCalibration(out);
  put_track(out, cal);
  return (int)out.size();
}
// LD:57-83 with the reference's own classes (the loop itself is restated here: it lives inside a ROS callback)
REF_API int ref_ld_long(const double* slam, const double* enu, int n, int iterations, double* w_out, double* cal) {
  std::vector<COORDXYZT> s = make_track(slam, n), e = make_track(enu, n);
  WeightCoeCal wcca;
  std::vector<double> wc;
  wcca.ICPWeightCoeCal(s, wc);
  std::vector<COORDXYZT> pro;
  {
    trackCalibration tc(s, e, wc);
    tc.doICP();
    tc.doCalibration(pro);
  }
  for (int i = 1; i <= iterations; i++) {
    wc.clear();
    pro.reserve((size_t)n + 1);
    wcca.ICPWeightCoeCal(s, wc, e, pro);
    trackCalibration tc2(pro, e, wc);
    pro.clear();
    tc2.doICP();
    tc2.doCalibration(pro);
  }
  for (int i = 0; i < n; i++) w_out[i] = wc[i];
  put_track(pro, cal);
  return n;
}
