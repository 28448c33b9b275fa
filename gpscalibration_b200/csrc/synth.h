// Deterministic synthetic LiDAR sweep generator — workload generator shared by tests and bench.py.
// It is input data only: neither the oracle nor the CUDA path depends on it.  The reference ships no LiDAR fixture
// (data/bag_list.txt:1 is a placeholder), so tests and bench.py drive both the oracle and the CUDA path with
// sweeps ray-cast from this scene.  Output: packed float xyz in the SENSOR frame (x fwd, y left, z up), firing
// order (azimuth-major, ring-minor, clockwise like a Velodyne so that SR:340's -atan2 increases), no-return rays
// dropped, the sensor moving during the sweep (so LO:123-150's de-skew has real work to do).
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

namespace loamsynth {

struct Box { double lo[3], hi[3]; };
struct Cyl { double cx, cy, r, h; };

struct Scene {
  std::vector<Box> boxes;
  std::vector<Cyl> cyls;
};

inline uint64_t splitmix64(uint64_t& s) {
  uint64_t z = (s += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
inline double u01(uint64_t& s) { return (double)(splitmix64(s) >> 11) * (1.0 / 9007199254740992.0); }

// scene 0: ring road of radius 50 m around a 60x60 m block, inside a 160x160 m walled yard, with cars, kiosks,
// poles and buildings along the road.  scene 1: same plus a dense grid of tall blocks (more vertical structure).
inline Scene make_scene(int kind, uint64_t seed) {
  Scene sc;
  uint64_t s = seed * 0x2545F4914F6CDD1Dull + 12345;
  auto box = [&](double x0, double y0, double z0, double x1, double y1, double z1) { sc.boxes.push_back(Box{{x0, y0, z0}, {x1, y1, z1}}); };
  box(-30, -30, 0, 30, 30, 25);          // inner block
  box(-81, -81, 0, -80, 81, 30);         // yard walls
  box(80, -81, 0, 81, 81, 30);
  box(-81, -81, 0, 81, -80, 30);
  box(-81, 80, 0, 81, 81, 30);
  int nobj = 28;
  for (int k = 0; k < nobj; k++) {
    double a = 2 * M_PI * (k + 0.3 * u01(s)) / nobj;
    bool inner = (k & 1);
    double rad = inner ? 43.5 + 1.0 * u01(s) : 57.0 + 3.0 * u01(s);
    double cx = rad * std::cos(a), cy = rad * std::sin(a);
    if (k % 3 == 0) {  // pole
      sc.cyls.push_back(Cyl{cx, cy, 0.15 + 0.15 * u01(s), 6.0 + 2.0 * u01(s)});
    } else if (k % 3 == 1) {  // car-sized box
      double hx = 1.0 + 1.2 * u01(s), hy = 1.0 + 1.2 * u01(s);
      box(cx - hx, cy - hy, 0, cx + hx, cy + hy, 1.4 + 0.5 * u01(s));
    } else {  // kiosk
      double h = 1.2 + 0.8 * u01(s);
      box(cx - h, cy - h, 0, cx + h, cy + h, 3.0 + 1.5 * u01(s));
    }
  }
  for (int k = 0; k < 12; k++) {  // buildings between road and wall
    double a = 2 * M_PI * (k + 0.5) / 12;
    double cx = 69 * std::cos(a), cy = 69 * std::sin(a);
    cx = std::max(-74.0, std::min(74.0, cx * 1.05));
    cy = std::max(-74.0, std::min(74.0, cy * 1.05));
    double hx = 3 + 2 * u01(s), hy = 3 + 2 * u01(s);
    box(cx - hx, cy - hy, 0, cx + hx, cy + hy, 8 + 12 * u01(s));
  }
  for (int k = 0; k < 16; k++) {  // extra poles
    double a = 2 * M_PI * (k + 0.25 + 0.4 * u01(s)) / 16;
    double rad = (k & 1) ? 45.0 : 55.5;
    sc.cyls.push_back(Cyl{rad * std::cos(a), rad * std::sin(a), 0.12 + 0.1 * u01(s), 5.0 + 3.0 * u01(s)});
  }
  if (kind == 1) {
    for (int i = -3; i <= 3; i++)
      for (int j = -3; j <= 3; j++) {
        double cx = 21.0 * i, cy = 21.0 * j;
        double r = std::sqrt(cx * cx + cy * cy);
        if (r < 36 || (r > 44 && r < 62) || std::fabs(cx) > 72 || std::fabs(cy) > 72) continue;
        box(cx - 4, cy - 4, 0, cx + 4, cy + 4, 10 + 15 * u01(s));
      }
  }
  return sc;
}

struct SensorModel {
  int n_rings, n_cols;
  std::vector<double> elev_deg;
};
// kind 0: 16 rings at the reference's own ring-table angles (SR:303-318) x 1800 columns ("VLP-16-shaped")
// kind 1: true VLP-16 angles (-15..+15 step 2) x 1800 -> exercises dropped beams / empty rings
// kind 2: 64 rings -24.8..+2.0 deg (0.4254 step) x 1875 columns ("HDL-64-shaped")
inline SensorModel make_sensor(int kind) {
  SensorModel m;
  if (kind == 0) {
    m.n_rings = 16; m.n_cols = 1800;
    m.elev_deg = {-15, -13, -11, -9, -7, -5, -4, -3, -2, -1, 0, 1, 3, 5, 7, 9};
  } else if (kind == 1) {
    m.n_rings = 16; m.n_cols = 1800;
    for (int r = 0; r < 16; r++) m.elev_deg.push_back(-15.0 + 2.0 * r);
  } else {
    m.n_rings = 64; m.n_cols = 1875;
    for (int r = 0; r < 64; r++) m.elev_deg.push_back(-24.8 + 26.8 * r / 63.0);
  }
  return m;
}

struct Pose { double x, y, z, yaw, pitch, roll; };
// constant 1.0 m/sweep on a 50 m circle (yaw rate 0.02 rad/sweep) with a little sway, bounce, pitch and roll
inline Pose trajectory(double t) {
  const double Rc = 50.0, w = 0.02;
  double th = w * t;
  double sway = 0.05 * std::sin(0.7 * t);
  Pose p;
  p.x = (Rc + sway) * std::sin(th);
  p.y = -(Rc + sway) * std::cos(th);
  p.z = 1.8 + 0.02 * std::sin(0.3 * t);
  p.yaw = th;
  p.pitch = 0.004 * std::sin(0.4 * t + 1.0);
  p.roll = 0.005 * std::sin(0.5 * t);
  return p;
}

inline void rot_zyx(const Pose& p, const double* v, double* o) {  // R = Rz(yaw) Ry(pitch) Rx(roll)
  double cr = std::cos(p.roll), sr = std::sin(p.roll), cp = std::cos(p.pitch), sp = std::sin(p.pitch), cy = std::cos(p.yaw), sy = std::sin(p.yaw);
  double x1 = v[0], y1 = cr * v[1] - sr * v[2], z1 = sr * v[1] + cr * v[2];
  double x2 = cp * x1 + sp * z1, y2 = y1, z2 = -sp * x1 + cp * z1;
  o[0] = cy * x2 - sy * y2;
  o[1] = sy * x2 + cy * y2;
  o[2] = z2;
}

inline double ray_scene(const Scene& sc, const double* o, const double* d, double tmax) {
  double best = tmax;
  if (d[2] < -1e-12) {  // ground z = 0
    double t = -o[2] / d[2];
    if (t > 0.05 && t < best) best = t;
  }
  for (const Box& b : sc.boxes) {
    double t0 = 0.05, t1 = best;
    bool hit = true;
    for (int a = 0; a < 3 && hit; a++) {
      if (std::fabs(d[a]) < 1e-12) {
        if (o[a] < b.lo[a] || o[a] > b.hi[a]) hit = false;
      } else {
        double inv = 1.0 / d[a];
        double ta = (b.lo[a] - o[a]) * inv, tb = (b.hi[a] - o[a]) * inv;
        if (ta > tb) std::swap(ta, tb);
        if (ta > t0) t0 = ta;
        if (tb < t1) t1 = tb;
        if (t0 > t1) hit = false;
      }
    }
    if (hit && t0 < best) best = t0;
  }
  for (const Cyl& c : sc.cyls) {
    double ox = o[0] - c.cx, oy = o[1] - c.cy;
    double A = d[0] * d[0] + d[1] * d[1];
    if (A < 1e-14) continue;
    double B = ox * d[0] + oy * d[1];
    double C = ox * ox + oy * oy - c.r * c.r;
    double disc = B * B - A * C;
    if (disc < 0) continue;
    double t = (-B - std::sqrt(disc)) / A;
    if (t > 0.05 && t < best) {
      double z = o[2] + t * d[2];
      if (z >= 0 && z <= c.h) best = t;
    }
  }
  return best;
}

// Returns the number of points written (<= n_rings * n_cols).  xyz must hold 3 * n_rings * n_cols floats.
inline int synth_sweep(const Scene& sc, const SensorModel& sm, uint64_t seed, int sweep_id, double t_offset, float* xyz,
                       double range_sigma = 0.02, double max_range = 100.0) {
  int n = 0;
  std::vector<double> ce(sm.n_rings), se(sm.n_rings);
  for (int r = 0; r < sm.n_rings; r++) {
    ce[r] = std::cos(sm.elev_deg[r] * M_PI / 180.0);
    se[r] = std::sin(sm.elev_deg[r] * M_PI / 180.0);
  }
  for (int c = 0; c < sm.n_cols; c++) {
    double tau = (double)c / sm.n_cols;
    Pose p = trajectory(t_offset + sweep_id + tau);
    double phi = -2 * M_PI * tau;
    double cph = std::cos(phi), sph = std::sin(phi);
    double o[3] = {p.x, p.y, p.z};
    for (int r = 0; r < sm.n_rings; r++) {
      double ds[3] = {ce[r] * cph, ce[r] * sph, se[r]};
      double dw[3];
      rot_zyx(p, ds, dw);
      double t = ray_scene(sc, o, dw, max_range);
      if (!(t < max_range)) continue;
      uint64_t s = seed ^ ((uint64_t)(uint32_t)sweep_id << 32) ^ (uint64_t)(c * sm.n_rings + r) * 0x9E3779B97F4A7C15ull;
      double u1 = u01(s), u2 = u01(s);
      if (u1 < 1e-300) u1 = 1e-300;
      double g = std::sqrt(-2.0 * std::log(u1)) * std::cos(2 * M_PI * u2);
      double rr = t + range_sigma * g;
      xyz[3 * n + 0] = (float)(ds[0] * rr);
      xyz[3 * n + 1] = (float)(ds[1] * rr);
      xyz[3 * n + 2] = (float)(ds[2] * rr);
      n++;
    }
  }
  return n;
}

}  // namespace loamsynth
