// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_linalg.h header).
//
// Restatement of the PCL 1.8.0 (+FLANN) pieces the reference's hot path calls.  PCL/FLANN are un-vendored
// third-party dependencies (PCL pinned to 1.8.0 by install/install_u1604_basic.sh:32; FLANN unpinned), so these
// follow the published algorithms and the reference's own call sites:
//   pcl::KdTreeFLANN::nearestKSearch  (LO:603,758; LM:760,867) -> knn_brute / KdTree::knn
//        exact kNN, L2_Simple accumulation ((dx*dx)+(dy*dy))+(dz*dz) in fp32, results ascending.
//        TIE RULE (ours): total order (d2 ascending, index ascending).
//   pcl::VoxelGrid<PointXYZI>::filter (SR:677-683; LM:736-744,1061-1079,1092-1094) -> voxel_grid
//        in-cell summation order fixed to ascending input index (PCL's std::sort is unstable).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <vector>

namespace orc {

struct P4 {
  float x, y, z, i;
};
typedef std::vector<P4> Cloud;

inline float sqdist(const P4& a, const P4& b) {
  float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
  return ((dx * dx) + (dy * dy)) + (dz * dz);
}

struct Nbr {
  float d2;
  int idx;
};
inline bool nbr_less(const Nbr& a, const Nbr& b) { return a.d2 < b.d2 || (a.d2 == b.d2 && a.idx < b.idx); }

// k best of `cloud` for query q; out[0..k') ascending by (d2, idx); returns k' = min(k, size).
inline int knn_brute(const Cloud& cloud, const P4& q, int k, Nbr* out) {
  int found = 0;
  for (int j = 0; j < (int)cloud.size(); j++) {
    Nbr c = {sqdist(cloud[j], q), j};
    if (found == k && !nbr_less(c, out[k - 1])) continue;
    int pos = found < k ? found : k - 1;
    while (pos > 0 && nbr_less(c, out[pos - 1])) {
      out[pos] = out[pos - 1];
      pos--;
    }
    out[pos] = c;
    if (found < k) found++;
  }
  return found;
}

// Exact kd-tree (median split on the widest axis, leaf size 15 like PCL's KDTreeSingleIndexParams(15)).
// Produces exactly knn_brute's answer (asserted in tests); only the visiting order differs.
class KdTree {
 public:
  void build(const Cloud& c) {
    cloud_ = &c;
    int n = (int)c.size();
    perm_.resize(n);
    for (int i = 0; i < n; i++) perm_[i] = i;
    nodes_.clear();
    if (n > 0) {
      nodes_.reserve(2 * (n / 8 + 1));
      build_rec(0, n);
    }
  }
  int size() const { return (int)perm_.size(); }
  int knn(const P4& q, int k, Nbr* out) const {
    int found = 0;
    if (!nodes_.empty()) search(0, q, k, out, found);
    return found;
  }

 private:
  struct Node {
    int lo, hi;       // range in perm_
    int left, right;  // children (-1 for leaf)
    int dim;
    float split;
  };
  const Cloud* cloud_ = nullptr;
  std::vector<int> perm_;
  std::vector<Node> nodes_;
  static float coord(const P4& p, int d) { return d == 0 ? p.x : (d == 1 ? p.y : p.z); }

  int build_rec(int lo, int hi) {
    int id = (int)nodes_.size();
    nodes_.push_back(Node{lo, hi, -1, -1, 0, 0.f});
    if (hi - lo <= 15) return id;
    float mn[3] = {1e30f, 1e30f, 1e30f}, mx[3] = {-1e30f, -1e30f, -1e30f};
    for (int i = lo; i < hi; i++) {
      const P4& p = (*cloud_)[perm_[i]];
      for (int d = 0; d < 3; d++) {
        float v = coord(p, d);
        mn[d] = std::min(mn[d], v);
        mx[d] = std::max(mx[d], v);
      }
    }
    int dim = 0;
    for (int d = 1; d < 3; d++)
      if (mx[d] - mn[d] > mx[dim] - mn[dim]) dim = d;
    if (!(mx[dim] > mn[dim])) return id;  // all coincident: keep as leaf
    int mid = (lo + hi) / 2;
    const Cloud& c = *cloud_;
    std::nth_element(perm_.begin() + lo, perm_.begin() + mid, perm_.begin() + hi, [&](int a, int b) {
      float va = coord(c[a], dim), vb = coord(c[b], dim);
      return va < vb || (va == vb && a < b);
    });
    float split = coord(c[perm_[mid]], dim);
    int l = build_rec(lo, mid);
    int r = build_rec(mid, hi);
    nodes_[id].left = l;
    nodes_[id].right = r;
    nodes_[id].dim = dim;
    nodes_[id].split = split;
    return id;
  }

  void offer(const Nbr& c, int k, Nbr* out, int& found) const {
    if (found == k && !nbr_less(c, out[k - 1])) return;
    int pos = found < k ? found : k - 1;
    while (pos > 0 && nbr_less(c, out[pos - 1])) {
      out[pos] = out[pos - 1];
      pos--;
    }
    out[pos] = c;
    if (found < k) found++;
  }

  void search(int id, const P4& q, int k, Nbr* out, int& found) const {
    const Node& nd = nodes_[id];
    if (nd.left < 0) {
      for (int i = nd.lo; i < nd.hi; i++) {
        int j = perm_[i];
        offer(Nbr{sqdist((*cloud_)[j], q), j}, k, out, found);
      }
      return;
    }
    float diff = coord(q, nd.dim) - nd.split;
    int near = diff < 0.f ? nd.left : nd.right;
    int far = diff < 0.f ? nd.right : nd.left;
    search(near, q, k, out, found);
    // every point on the far side has |d_dim| >= |diff| hence (fp-monotone) d2 >= diff*diff: prune only if strictly worse
    float pd = diff * diff;
    if (found < k || !(pd > out[k - 1].d2)) search(far, q, k, out, found);
  }
};

// pcl::VoxelGrid<PointXYZI>::applyFilter, PCL 1.8.0 defaults (downsample_all_data, min_points_per_voxel = 0).
// Returns false when PCL would refuse (index overflow) and hand the input back unfiltered.
inline bool voxel_grid(const Cloud& in, float leaf, Cloud& out) {
  out.clear();
  if (in.empty()) return true;
  float inv = 1.0f / leaf;
  float mnx = std::numeric_limits<float>::max(), mny = mnx, mnz = mnx;
  float mxx = -std::numeric_limits<float>::max(), mxy = mxx, mxz = mxx;
  for (const P4& p : in) {
    mnx = std::min(mnx, p.x); mny = std::min(mny, p.y); mnz = std::min(mnz, p.z);
    mxx = std::max(mxx, p.x); mxy = std::max(mxy, p.y); mxz = std::max(mxz, p.z);
  }
  int64_t dx = (int64_t)((mxx - mnx) * inv) + 1;
  int64_t dy = (int64_t)((mxy - mny) * inv) + 1;
  int64_t dz = (int64_t)((mxz - mnz) * inv) + 1;
  if (dx * dy * dz > (int64_t)std::numeric_limits<int32_t>::max()) {
    out = in;
    return false;
  }
  int minb[3] = {(int)floorf(mnx * inv), (int)floorf(mny * inv), (int)floorf(mnz * inv)};
  int maxb[3] = {(int)floorf(mxx * inv), (int)floorf(mxy * inv), (int)floorf(mxz * inv)};
  int divx = maxb[0] - minb[0] + 1, divy = maxb[1] - minb[1] + 1;
  int mul1 = divx, mul2 = divx * divy;
  struct CI {
    int cell, idx;
  };
  std::vector<CI> v(in.size());
  for (int n = 0; n < (int)in.size(); n++) {
    const P4& p = in[n];
    int i0 = (int)(floorf(p.x * inv) - (float)minb[0]);
    int i1 = (int)(floorf(p.y * inv) - (float)minb[1]);
    int i2 = (int)(floorf(p.z * inv) - (float)minb[2]);
    v[n].cell = i0 + i1 * mul1 + i2 * mul2;
    v[n].idx = n;
  }
  std::sort(v.begin(), v.end(), [](const CI& a, const CI& b) { return a.cell < b.cell || (a.cell == b.cell && a.idx < b.idx); });
  size_t a = 0;
  while (a < v.size()) {
    size_t b = a;
    float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
    while (b < v.size() && v[b].cell == v[a].cell) {
      const P4& p = in[v[b].idx];
      sx = sx + p.x; sy = sy + p.y; sz = sz + p.z; si = si + p.i;
      b++;
    }
    float n = (float)(b - a);
    out.push_back(P4{sx / n, sy / n, sz / n, si / n});
    a = b;
  }
  return true;
}

}  // namespace orc
