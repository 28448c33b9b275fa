// libloamgpu C ABI (include/loamgpu.h): the handle, the node-level state machines that replace the bodies of the
// reference's three LOAM nodes, and the stage-level entry points the parity tests call.
// Host logic restated here (the reference keeps it on the CPU and so do we): the Gauss-Newton solve and degeneracy
// projection LO:975-1004 / LM:968-997, convergence tests LO:1017-1028 / LM:1006-1017, pose accumulation LO:1035-1064,
// transformAssociateToMap / transformUpdate LM:120-242, the rolling cube grid LM:489-715 (descriptor moves only) and
// the reset protocol LO:411-415,519-563 / LM:316-319,434-461.
// There is NO CPU fallback: every compute step below is a kernel launch; without a CUDA device loam_create fails.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <immintrin.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "lg_extract.h"
#include "lg_host.h"
#include "lg_linalg.cuh"
#include "lg_map.h"
#include "lg_odom.h"
#include "lg_reduce.cuh"
#include "lg_voxel.h"

static thread_local char g_cuda_err[512] = "";
thread_local LgProf* g_lg_prof = nullptr;
// Exchange buffers (loam_shard_export) of THIS process: a CUDA IPC handle cannot be opened by its exporter, so ranks
// that live in one process (several handles driven by host threads) are connected by pointer.
static std::mutex g_xchg_mutex;
static std::map<std::string, double*> g_xchg_local;
void lg_set_error(const char* msg, const char* file, int line) { snprintf(g_cuda_err, sizeof(g_cuda_err), "%s (%s:%d)", msg, file, line); }

#define LG_SYNC(h) do { (h)->syncs++; LG_CHECK(cudaStreamSynchronize((h)->st)); } while (0)
#define LG_D2H(h, dst, src, bytes) do { (h)->d2h_bytes += (long long)(bytes); LG_CHECK(cudaMemcpyAsync((dst), (src), (bytes), cudaMemcpyDeviceToHost, (h)->st)); } while (0)

namespace {

enum { HT_EXTRACT = 0, HT_ODOM_ITERS = 1, HT_ODOM_END = 2, HT_MAP_PREP = 3, HT_MAP_GRID = 4, HT_MAP_ITERS = 5, HT_MAP_INSERT = 6,
       HT_MAP_CUBEDS = 7, HT_MAP_REST = 8 };
struct HostTimer {
  double* acc;
  std::chrono::steady_clock::time_point t0;
  explicit HostTimer(double* a) : acc(a), t0(std::chrono::steady_clock::now()) {}
  void lap(double* next) {
    auto t1 = std::chrono::steady_clock::now();
    *acc += std::chrono::duration<double>(t1 - t0).count();
    acc = next;
    t0 = t1;
  }
  ~HostTimer() { *acc += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
};

struct Chunk {
  int off, n;
};
constexpr int CW = 21, CH = 11, CD = 21, CNUM = CW * CH * CD;  // LM:72-75

static SinCos3 host_sincos3(const float* T) {
  SinCos3 s;
  s.srx = sinf(T[0]); s.crx = cosf(T[0]);
  s.sry = sinf(T[1]); s.cry = cosf(T[1]);
  s.srz = sinf(T[2]); s.crz = cosf(T[2]);
  return s;
}

}  // namespace

struct loam_handle {
  loam_params prm;
  int device = 0;
  cudaStream_t st = nullptr;
  long long launches = 0;
  long long h2d_bytes = 0, d2h_bytes = 0, syncs = 0;
  unsigned long long mail_seq = 0;  // sequence number the reduction kernels publish into the mapped mailbox
  LgProf prof;
  double host_s[LOAM_HOST_SECTIONS] = {0};  // host wall-clock per section (diagnostics, loam_host_times)
  // Pipelined mode only: /laser_cloud_surround (LM:1081-1101) is an output nobody downstream in the path waits for, so the
  // mapping stage only ENQUEUES the gather of the cubes on the output handle's stream (`aux`) and goes on with the next
  // sweep; the voxel grid and its count read-back run on the pipeline's output thread.  ev_map_done: this handle's map
  // kernels of the run are finished; ev_aux_read: the gather has read the arena (awaited before the arena is written again).
  // The per-cube voxel grids at the end of a mapping run (LM:1061-1079) only change what the NEXT run gathers, so their
  // counts are not waited for: the read-back is enqueued, and the cube descriptors are brought up to date when the next run
  // starts (by then the kernels are long finished) -- unless this run still needs them for the surround cloud.
  struct PendingDS {
    bool active = false, merged = false;
    int nseg = 0, Mtot = 0, max_n = 0;
    std::vector<int> validInd, seg_off;
    std::vector<float> leaf;
    std::vector<CopyEnt> ents;  // segment-major input layout, for the (rare) fall-back from the merge path to the full sort
    cudaEvent_t ev = nullptr;
  } pds;
  // the voxel-grid chain itself runs on a second stream behind the run's insert / sort (ev_pre_ds), so that the NEXT run's
  // stack transform and stack voxel grid (which need neither the cubes nor this chain's scratch) overlap with it
  cudaStream_t st2 = nullptr;
  cudaEvent_t ev_pre_ds = nullptr;
  DevBuf batch_tab;        // argument table of loam_extract_batch (first handle of the batch)
  int* h_ints2 = nullptr;  // pinned: [start | end | merge flag] of the pending read-back
  loam_handle* aux = nullptr;
  cudaEvent_t ev_map_done = nullptr, ev_aux_read = nullptr;
  bool aux_read_pending = false;
  bool aux_reserved = false;  // the pipeline holds the output handle's scratch for this run (else the surround cloud is made in line)
  int aux_ns = 0;  // points gathered for the output thread
  // pinned host staging
  double* h_mail = nullptr;  // mapped: 28 doubles written by the reduction kernels
  double* d_mail = nullptr;
  int* h_ints = nullptr;     // pinned scratch for small device -> host reads
  static constexpr int H_INTS = 8192;

  // ---- scanRegistration
  SrParams srp;
  SrWs sr;
  DevBuf xyz_in;
  loam_counts counts = {0, 0, 0, 0, 0};
  bool have_features = false;
  float imu[12] = {0};
  // the node's IMU state (SR:76-110): ring of integrated messages (host copy + device copy), the device-resident carry of the
  // per-point de-skew; inactive until the first loam_imu_push (imuPointerLast == -1, SR:364)
  lgh::ImuHost imu_host;
  SrImuRing imu_ring = {};
  DevBuf d_imu_ring, d_imu_carry;
  SrImuCarry imu_carry = {};  // host mirror of the device carry after the last sweep
  bool imu_ring_dirty = false;
  // current message set for odometry (device pointers; either sr.* or the explicit test buffers)
  const float4 *cur_sharp = nullptr, *cur_less_sharp = nullptr, *cur_flat = nullptr, *cur_less_flat = nullptr, *cur_full = nullptr;
  DevBuf t_sharp, t_flat;  // loam_odom_set_inputs

  // ---- laserOdometry state (the reference's file-scope / main()-scope variables)
  OdomWs od;
  bool lo_inited = false;          // systemInited LO:53
  int frameCount = 1;              // LO:495 (= skipFrameNum)
  float T[6] = {0}, Tsum[6] = {0};  // transformation / transformationSum LO:111-112
  LgGNState lo_gn;                 // matP / isDegenerate LO:489-492
  DevBuf xyz_packed, wire;
  // sharded map with the fused all-reduce (loam_shard_*): exchange buffer, peers' mappings, iteration counter
  double* xchg = nullptr;
  PeerXchg px{};
  bool px_connected = false;
  bool px_local[LG_MAX_PEERS] = {};
  DevBuf corner_last, surf_last, corner_new, surf_new, fullres3;
  int n_corner_last = 0, n_surf_last = 0, n_fullres3 = 0;
  int cornerLastNum = 0, surfLastNum = 0;  // LO:98-99 (gate values, lag one sweep behind after init)

  // ---- transformMaintenance state (TM:77-81, 99-100)
  float tmSum[6] = {0}, tmIncre[6] = {0}, tmMapped[6] = {0}, tmBef[6] = {0}, tmAft[6] = {0};
  double tmPre[4] = {0, 0, 0, 0}, tmTmp[4] = {0, 0, 0, 0};

  // ---- laserMapping state
  bool lm_inited = false;  // systemInited LM:49
  int mapFrameCount = 4;   // LM:419
  int cenW = 10, cenH = 5, cenD = 10;
  float mTsum[6] = {0}, Tincre[6] = {0}, Ttobe[6] = {0}, Tbef[6] = {0}, Taft[6] = {0};
  LgGNState lm_gn;
  std::vector<std::vector<Chunk>> cubeC, cubeS;  // laserCloudCornerArray / laserCloudSurfArray LM:98-99 as arena chunks
  DevBuf arena, arena2;
  bool use_merge_path = true;   // voxel-grid valid cubes by merging sorted old clouds with the sorted new points
  long long merge_fallbacks = 0;  // times the merge path's sortedness check failed and the full sort ran instead
  size_t bump = 0;  // in points
  DevBuf stack2_c, stack2_s, stack_c, stack_s, map_c, map_s;
  int n_stack_c = 0, n_stack_s = 0, n_map_c = 0, n_map_s = 0;
  CsrWs csr;     // cell-sorted grids over the gathered corner / surf map (replace the kd-trees, LM:750-751)
  MapGnWs gn;    // workspace of the fused Gauss-Newton kernel
  bool grids_valid = false;
  bool nbr_valid = false;  // gn.nbr holds the neighbours of the last stage-level iteration
  float slab_lo = -INFINITY, slab_hi = INFINITY;  // sharded map: this rank evaluates queries whose map-frame x is in [lo, hi)
  DevBuf d_ents, d_segs, d_ints, d_seg_off, d_seg_leaf, d_out_se;
  VoxBigWs vb;
  DevBuf ds_in, ins_sel, ins_sorted, d_runs;
  DevBuf surround, registered, vg_in, vg_out, vs_staging, vs_counts;
  int n_surround = 0, n_registered = 0;
};

namespace {

int upload_on(loam_handle* h, cudaStream_t st, DevBuf& dst, const void* src, size_t bytes) {
  h->h2d_bytes += (long long)bytes;
  LG_CHECK(dst.ensure(bytes + 16, st));
  if (bytes) LG_CHECK(cudaMemcpyAsync(dst.p, src, bytes, cudaMemcpyHostToDevice, st));
  return LOAM_OK;
}
int upload(loam_handle* h, DevBuf& dst, const void* src, size_t bytes) { return upload_on(h, h->st, dst, src, bytes); }

// ------------------------------------------------------------------------------------------------ map storage
size_t live_points(const loam_handle* h) {
  size_t s = 0;
  for (auto& v : h->cubeC)
    for (auto& c : v) s += c.n;
  for (auto& v : h->cubeS)
    for (auto& c : v) s += c.n;
  return s;
}

// Makes room for `need` more points at the bump pointer.  Voxel-gridding a cube writes its new cloud at the bump
// pointer and turns the old chunks into garbage; when the arena is full the live chunks are compacted into the
// second (ping-pong) arena — no cudaMalloc / cudaFree on the steady-state path.
int arena_reserve(loam_handle* h, size_t need) {
  size_t cap = h->arena.cap / 16;
  if (h->bump + need <= cap) return LOAM_OK;
  size_t live = live_points(h);
  size_t want = std::max<size_t>(8 * (live + need), (size_t)std::max(h->prm.max_map_points, 1 << 20) * 4);
  if (h->arena2.cap / 16 < live + need || h->arena2.cap < h->arena.cap) {
    LG_CHECK(h->arena2.ensure(std::max(want * 16, h->arena.cap), h->st));
  }
  std::vector<CopyEnt> ents;
  size_t off = 0;
  int max_n = 0;
  const float4* old = h->arena.as<float4>();
  for (auto* arr : {&h->cubeC, &h->cubeS})
    for (auto& v : *arr)
      for (auto& c : v) {
        if (c.n > 0) {
          ents.push_back(CopyEnt{old + c.off, c.n, (int)off});
          max_n = std::max(max_n, c.n);
        }
        c.off = (int)off;
        off += c.n;
      }
  if (!ents.empty()) {
    int rc = upload(h, h->d_ents, ents.data(), ents.size() * sizeof(CopyEnt));
    if (rc) return rc;
    rc = lg_gather(h->d_ents.as<CopyEnt>(), (int)ents.size(), max_n, h->arena2.as<float4>(), h->st, &h->launches);
    if (rc) return rc;
  }
  std::swap(h->arena, h->arena2);
  h->bump = off;
  return LOAM_OK;
}

void map_reset(loam_handle* h) {  // LM:434-461
  h->lm_gn = LgGNState();
  for (auto& v : h->cubeC) v.clear();
  for (auto& v : h->cubeS) v.clear();
  h->bump = 0;
  h->mapFrameCount = 4;
  h->cenW = 10; h->cenH = 5; h->cenD = 10;
  for (int i = 0; i < 6; i++) h->Tincre[i] = h->Ttobe[i] = h->Tbef[i] = h->Taft[i] = 0.f;
  h->grids_valid = false;
}

// Move every cube one step along `axis` (+1: towards higher index, the last slab re-enters at 0 emptied), LM:497-657.
void cube_shift(loam_handle* h, int axis, int dir) {
  int dims[3] = {CW, CH, CD};
  int n = dims[axis], a1 = (axis + 1) % 3, a2 = (axis + 2) % 3;
  for (auto* arr : {&h->cubeC, &h->cubeS})
    for (int u = 0; u < dims[a1]; u++)
      for (int v = 0; v < dims[a2]; v++) {
        auto idx = [&](int t) {
          int c[3];
          c[axis] = t; c[a1] = u; c[a2] = v;
          return c[0] + CW * c[1] + CW * CH * c[2];
        };
        if (dir > 0) {
          for (int t = n - 1; t >= 1; t--) (*arr)[idx(t)].swap((*arr)[idx(t - 1)]);
          (*arr)[idx(0)].clear();
        } else {
          for (int t = 0; t < n - 1; t++) (*arr)[idx(t)].swap((*arr)[idx(t + 1)]);
          (*arr)[idx(n - 1)].clear();
        }
      }
}

// Gathers the chunks of the listed cubes (in list order) into dst; returns the point count through *n_out.
int gather_cubes(loam_handle* h, const std::vector<int>& cubes, bool corner, bool surf, DevBuf& dst, int* n_out) {
  std::vector<CopyEnt> ents;
  const float4* ar = h->arena.as<float4>();
  int off = 0, max_n = 0;
  for (int ind : cubes) {
    if (corner)
      for (auto& c : h->cubeC[ind])
        if (c.n > 0) { ents.push_back(CopyEnt{ar + c.off, c.n, off}); off += c.n; max_n = std::max(max_n, c.n); }
    if (surf)
      for (auto& c : h->cubeS[ind])
        if (c.n > 0) { ents.push_back(CopyEnt{ar + c.off, c.n, off}); off += c.n; max_n = std::max(max_n, c.n); }
  }
  *n_out = off;
  LG_CHECK(dst.ensure((size_t)(off + 16) * 16, h->st));
  if (ents.empty()) return LOAM_OK;
  int rc = upload(h, h->d_ents, ents.data(), ents.size() * sizeof(CopyEnt));
  if (rc) return rc;
  return lg_gather(h->d_ents.as<CopyEnt>(), (int)ents.size(), max_n, dst.as<float4>(), h->st, &h->launches);
}

// Voxel grid of a handful of device-resident clouds (segments); counts come back through pinned memory.
// Small path when every segment fits the shared-memory sort, big path otherwise.
int voxel_segments(loam_handle* h, const std::vector<VoxSegD>& segs_in, std::vector<int>& counts) {
  const int nseg = (int)segs_in.size();
  counts.assign(nseg, 0);
  int max_n = 0;
  for (auto& s : segs_in) max_n = std::max(max_n, s.n);
  if (max_n == 0) return LOAM_OK;
  LG_CHECK(h->d_ints.ensure(4096, h->st));
  int* d_counts = h->d_ints.as<int>();
  if (max_n <= 16384) {
    std::vector<VoxSegD> segs = segs_in;
    for (int i = 0; i < nseg; i++) segs[i].out_count = d_counts + 1 + i;
    int rc = upload(h, h->d_segs, segs.data(), nseg * sizeof(VoxSegD));
    if (rc) return rc;
    double units = 0;
    for (auto& sg : segs) units += sg.n;
    LgProfScope prof_scope(LGK_VOXEL, h->st, units);
    bool split = max_n > 4096;
    for (auto& sg : segs) split = split && sg.valid == nullptr;
    if (split) {  // sweep-sized stacks: 32 CTAs per segment instead of one
      LG_CHECK(cudaMemsetAsync(d_counts, 0, 4, h->st));
      rc = lg_vox_split(h->vs_staging, h->vs_counts, h->d_segs.as<VoxSegD>(), nseg, d_counts, h->st, &h->launches);
      if (rc) return rc;
      LG_D2H(h, h->h_ints, d_counts, (nseg + 1) * 4);
      LG_SYNC(h);
      if (h->h_ints[0] == 0) {
        for (int i = 0; i < nseg; i++) counts[i] = h->h_ints[1 + i];
        return LOAM_OK;
      }
    }
    rc = lg_vox_small(h->d_segs.as<VoxSegD>(), nseg, max_n, d_counts, h->st, &h->launches);
    if (rc) return rc;
    LG_D2H(h, h->h_ints, d_counts + 1, nseg * 4);
    LG_SYNC(h);
    for (int i = 0; i < nseg; i++) counts[i] = h->h_ints[i];
    return LOAM_OK;
  }
  // big path: stage the segments contiguously (it shares the sort workspace and the staging buffers with the per-cube voxel
  // grids of the previous mapping run, which may still be running on the second stream)
  if (h->pds.active) LG_CHECK(cudaStreamWaitEvent(h->st, h->pds.ev, 0));
  std::vector<CopyEnt> ents;
  std::vector<int> seg_off(nseg + 1, 0);
  std::vector<float> leaf(nseg);
  for (int i = 0; i < nseg; i++) {
    if (segs_in[i].valid) return LOAM_EINVAL;
    if (segs_in[i].n > 0) ents.push_back(CopyEnt{segs_in[i].in, segs_in[i].n, seg_off[i]});
    seg_off[i + 1] = seg_off[i] + segs_in[i].n;
    leaf[i] = segs_in[i].leaf;
  }
  const int M = seg_off[nseg];
  LG_CHECK(h->ds_in.ensure((size_t)(M + 16) * 16, h->st));
  LG_CHECK(h->vg_out.ensure((size_t)(M + 16) * 16, h->st));
  int rc = upload(h, h->d_ents, ents.data(), ents.size() * sizeof(CopyEnt));
  if (rc) return rc;
  rc = lg_gather(h->d_ents.as<CopyEnt>(), (int)ents.size(), max_n, h->ds_in.as<float4>(), h->st, &h->launches);
  if (rc) return rc;
  rc = upload(h, h->d_seg_off, seg_off.data(), (nseg + 1) * 4);
  if (rc) return rc;
  rc = upload(h, h->d_seg_leaf, leaf.data(), nseg * 4);
  if (rc) return rc;
  LG_CHECK(h->d_out_se.ensure((size_t)nseg * 8 + 16, h->st));
  int* d_start = h->d_out_se.as<int>();
  int* d_end = d_start + nseg;
  rc = lg_vox_big(h->vb, h->ds_in.as<float4>(), h->d_seg_off.as<int>(), h->d_seg_leaf.as<float>(), nseg, M, h->vg_out.as<float4>(), d_start,
                  d_end, h->st, &h->launches);
  if (rc) return rc;
  if (2 * nseg > loam_handle::H_INTS) return LOAM_ENOSPC;
  LG_D2H(h, h->h_ints, d_start, nseg * 8);
  LG_SYNC(h);
  for (int i = 0; i < nseg; i++) {
    int s = h->h_ints[i], e = h->h_ints[nseg + i];
    counts[i] = e - s;
    if (counts[i] > 0)
      LG_CHECK(cudaMemcpyAsync(segs_in[i].out, h->vg_out.as<float4>() + s, (size_t)counts[i] * 16, cudaMemcpyDeviceToDevice, h->st));
  }
  return LOAM_OK;
}

// h->h_ints holds the head of the extraction's meta array (read back by the caller): virtual rings, errors, counts
int sr_counts_tail(loam_handle* h, loam_counts* out) {
  if (h->h_ints[SRM_VIRTUAL]) {
    // A ring whose scanStartInd was never written (empty rings: true VLP-16 angles leave rings 6 / 8 / 10 of the
    // reference's table empty): it spans [0, scanEndInd) and is replayed after the rings before it, serially (SR:480-490).
    const int R = h->prm.n_scans, n = h->h_ints[SRM_N_FULL];
    LG_D2H(h, h->h_ints, h->sr.meta.p, SRM_SIZE * 4);
    LG_SYNC(h);
    std::vector<int> vr, vE;
    size_t stage = 0;
    for (int r = 1; r < R; r++) {
      const int S = h->h_ints[SRM_SCAN_START + r], E = (r == R - 1) ? n - 5 : h->h_ints[SRM_SCAN_END + r];
      if (S == 0 && E > 0) {
        vr.push_back(r);
        vE.push_back(E);
        stage += (size_t)E;
      }
    }
    int rc = lg_extract_virtual_launch(h->sr, h->srp, n, stage, h->st, &h->launches);
    if (rc) return rc;
    LG_D2H(h, h->h_ints, h->sr.lf_meta.p, (size_t)2 * R * 4);
    LG_SYNC(h);
    std::vector<VoxSegD> segs(vr.size());
    std::vector<int> offs(vr.size());
    for (size_t i = 0; i < vr.size(); i++) {
      offs[i] = h->h_ints[2 * vr[i]];
      segs[i] = VoxSegD{h->sr.lf_stage.as<float4>() + offs[i], nullptr, h->sr.lf_vout.as<float4>() + offs[i], nullptr, h->h_ints[2 * vr[i] + 1], 0.2f};
    }
    std::vector<int> cnt;
    rc = voxel_segments(h, segs, cnt);  // SR:677-683, one VoxelGrid per ring
    if (rc) return rc;
    for (size_t i = 0; i < vr.size(); i++) {  // patch the ring's voxel job: concat reads segs[r].out and meta[SRM_LF_CNT + r]
      VoxSegD sg = segs[i];
      sg.out_count = h->sr.meta.as<int>() + SRM_LF_CNT + vr[i];
      h->h2d_bytes += sizeof(VoxSegD) + 4;
      LG_CHECK(cudaMemcpyAsync(h->sr.segs.as<VoxSegD>() + vr[i], &sg, sizeof(VoxSegD), cudaMemcpyHostToDevice, h->st));
      LG_CHECK(cudaMemcpyAsync(h->sr.meta.as<int>() + SRM_LF_CNT + vr[i], &cnt[i], 4, cudaMemcpyHostToDevice, h->st));
    }
    LG_CHECK(cudaStreamSynchronize(h->st));  // sg / cnt are stack and vector storage
    rc = lg_extract_finish_launch(h->sr, h->srp, h->st, &h->launches);
    if (rc) return rc;
    LG_D2H(h, h->h_ints, h->sr.meta.p, SRM_HEAD * 4);
    LG_SYNC(h);
  }
  if (h->h_ints[SRM_ERR]) return LOAM_ENOSPC;
  if (h->h_ints[SRM_VOX_OVERFLOW]) return LOAM_ENOSPC;
  h->counts.n_full = h->h_ints[SRM_N_FULL];
  h->counts.n_sharp = h->h_ints[SRM_N_SHARP];
  h->counts.n_less_sharp = h->h_ints[SRM_N_LESS_SHARP];
  h->counts.n_flat = h->h_ints[SRM_N_FLAT];
  h->counts.n_less_flat = h->h_ints[SRM_N_LESS_FLAT];
  if (out) *out = h->counts;
  return LOAM_OK;
}
int read_sr_counts(loam_handle* h, loam_counts* out) {
  LG_D2H(h, h->h_ints, h->sr.meta.p, SRM_HEAD * 4);
  LG_SYNC(h);
  return sr_counts_tail(h, out);
}

// sensor_msgs/PointCloud2 payloads whose point_step is not a multiple of four (the Velodyne driver's 22-byte
// PointXYZIR: x y z @0, intensity @16, ring @20) cannot be read with aligned 4-byte loads: repack x y z first.
__global__ void unpack_xyz_kernel(const unsigned char* __restrict__ src, int n, int step, float* __restrict__ dst) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned char* p = src + (size_t)i * step;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    unsigned int v = (unsigned int)p[4 * k] | ((unsigned int)p[4 * k + 1] << 8) | ((unsigned int)p[4 * k + 2] << 16) | ((unsigned int)p[4 * k + 3] << 24);
    dst[3 * i + k] = __uint_as_float(v);
  }
}
// pcl::toROSMsg(pcl::PointCloud<pcl::PointXYZI>): point_step 32, x @0, y @4, z @8, padding 1.0f @12, intensity @16, zeros.
__global__ void pack_pcl32_kernel(const float4* __restrict__ src, int n, float4* __restrict__ dst) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = src[i];
  dst[2 * i] = make_float4(p.x, p.y, p.z, 1.0f);
  dst[2 * i + 1] = make_float4(p.w, 0.f, 0.f, 0.f);
}

// the five clouds of the sweep become the handle's current features (what the node publishes, SR:689-726)
void extract_publish(loam_handle* h) {
  h->cur_sharp = h->sr.sharp.as<float4>();
  h->cur_less_sharp = h->sr.less_sharp.as<float4>();
  h->cur_flat = h->sr.flat.as<float4>();
  h->cur_less_flat = h->sr.less_flat.as<float4>();
  h->cur_full = h->sr.full.as<float4>();
  h->have_features = true;
}

int extract_common(loam_handle* h, const float* d_xyz, int n, int stride_bytes, const float* imu_trans, loam_counts* out, double stamp = 0.0) {
  HostTimer ht(&h->host_s[HT_EXTRACT]);
  if (n < 0 || stride_bytes < 12) return LOAM_EINVAL;
  if (n > 0 && ((stride_bytes & 3) || ((uintptr_t)d_xyz & 3))) {  // unaligned wire layout: repack to 12-byte points
    LG_CHECK(h->xyz_packed.ensure((size_t)n * 12 + 16, h->st));
    unpack_xyz_kernel<<<lg_div_up(n, 256), 256, 0, h->st>>>((const unsigned char*)d_xyz, n, stride_bytes, h->xyz_packed.as<float>());
    h->launches++;
    LG_CHECK(cudaGetLastError());
    d_xyz = h->xyz_packed.as<float>();
    stride_bytes = 12;
  }
  for (int i = 0; i < 12; i++) h->imu[i] = imu_trans ? imu_trans[i] : 0.f;
  SrImuJob job;
  const bool with_imu = h->imu_host.last >= 0 && n > 0;  // SR:364: the IMU branch runs once a message has arrived
  if (with_imu) {
    if (!h->d_imu_carry.p) {
      LG_CHECK(h->d_imu_carry.ensure(sizeof(SrImuCarry), h->st));
      LG_CHECK(cudaMemsetAsync(h->d_imu_carry.p, 0, sizeof(SrImuCarry), h->st));
    }
    if (h->imu_ring_dirty || !h->d_imu_ring.p) {
      int rcu = upload(h, h->d_imu_ring, &h->imu_ring, sizeof(SrImuRing));  // pageable source: staged before the call returns
      if (rcu) return rcu;
      h->imu_ring_dirty = false;
    }
    job.ring = h->d_imu_ring.as<SrImuRing>();
    job.carry = h->d_imu_carry.as<SrImuCarry>();
    job.last = h->imu_host.last;
    job.time_scan = stamp;
  }
  int rc = lg_extract_launch(h->sr, h->srp, d_xyz, n, stride_bytes, h->st, &h->launches, with_imu ? &job : nullptr);
  if (rc) return rc;
  if (with_imu) LG_D2H(h, &h->imu_carry, h->d_imu_carry.p, sizeof(SrImuCarry));  // lands with the counts below
  rc = read_sr_counts(h, out);
  if (rc) return rc;
  if (with_imu) {  // /imu_trans as the node publishes it (SR:730-745); it overrides a caller-supplied imu_trans
    const SrImuCarry& c = h->imu_carry;
    const float tr[12] = {c.start[1], c.start[2], c.start[0], c.cur[1], c.cur[2], c.cur[0], c.shift_from_start[0], c.shift_from_start[1],
                          c.shift_from_start[2], c.velo_from_start[0], c.velo_from_start[1], c.velo_from_start[2]};
    for (int i = 0; i < 12; i++) h->imu[i] = tr[i];
  }
  extract_publish(h);
  return LOAM_OK;
}

// Waits until the reduction kernel has published sequence number h->mail_seq next to the 28 sums in the mapped pinned
// mailbox.  Spinning on host memory the GPU writes over PCIe costs ~2 us; cudaStreamSynchronize costs 10-20 us, and
// there is one such hand-over per host-visited Gauss-Newton iteration (every mapping iteration, the first odometry
// iteration of a sweep and once per device loop launch).
int mailbox_wait(loam_handle* h, cudaStream_t watch = nullptr) {  // watch: the stream the publishing kernel runs on (default: the handle's)
  if (!watch) watch = h->st;
  volatile unsigned long long* flag = (volatile unsigned long long*)(h->h_mail + 31);
  h->syncs++;
  for (long spin = 0;; spin++) {
    if (*flag == h->mail_seq) {
      std::atomic_thread_fence(std::memory_order_acquire);  // the 28 payload doubles are read after the flag
      return LOAM_OK;
    }
    _mm_pause();
    if ((spin & 0xffff) == 0xffff) {  // every ~65k polls make sure the stream has not died
      cudaError_t e = cudaStreamQuery(watch);
      if (e != cudaSuccess && e != cudaErrorNotReady) {
        lg_set_error(cudaGetErrorString(e), __FILE__, __LINE__);
        return LOAM_ECUDA;
      }
      if (e == cudaSuccess && *flag != h->mail_seq) {  // stream drained but no flag: should not happen
        if (*flag == h->mail_seq) {
          std::atomic_thread_fence(std::memory_order_acquire);
          return LOAM_OK;
        }
        lg_set_error("mailbox sequence never arrived", __FILE__, __LINE__);
        return LOAM_ECUDA;
      }
    }
  }
}

// One odometry iteration: launch, wait for the 28-double mailbox, unpack.
int odom_iter(loam_handle* h, int iter, const float* T, float* AtA, float* AtB, int* n_sel) {
  OdomT ot;
  for (int i = 0; i < 6; i++) ot.t[i] = T[i];
  SinCos3 sc = host_sincos3(T);
  int rc = lg_odom_iter_launch(h->od, ot, sc, iter, h->cur_sharp, h->counts.n_sharp, h->cur_flat, h->counts.n_flat,
                               h->corner_last.as<float4>(), h->n_corner_last, h->surf_last.as<float4>(), h->n_surf_last, h->d_mail,
                               ++h->mail_seq, h->st, &h->launches);
  if (rc) return rc;
  rc = mailbox_wait(h);
  if (rc) return rc;
  h->d2h_bytes += 28 * 8;
  lg_unpack28(h->h_mail, AtA, AtB, n_sel);
  return LOAM_OK;
}

// Enqueues iterations [it0, it1) of the scan-to-map Gauss-Newton loop (lg_mapgn.cu).  solve = 0: one pass, the 28 sums go
// to out_dev (device memory, no sequence word) or to the host mailbox; solve = 1: the loop runs on the device, the mailbox
// receives the pose, the iteration count and the sums of iteration 0.
int map_gn_enqueue(loam_handle* h, const float* T, int it0, int it1, int solve, double* out_dev, bool use_px) {
  MapGnArgs A;
  memset(&A, 0, sizeof(A));
  const SinCos3 sc = host_sincos3(T);
  const float scv[6] = {sc.srx, sc.crx, sc.sry, sc.cry, sc.srz, sc.crz};
  for (int i = 0; i < 6; i++) {
    A.T[i] = T[i];
    A.sc[i] = scv[i];
  }
  memcpy(A.matP, h->lm_gn.matP, sizeof(A.matP));
  A.degenerate = h->lm_gn.degenerate ? 1 : 0;
  A.it0 = it0;
  A.it1 = it1;
  A.solve = solve;
  A.cstack = h->stack_c.as<float4>();
  A.n_cs = h->n_stack_c;
  A.sstack = h->stack_s.as<float4>();
  A.n_ss = h->n_stack_s;
  A.gc = h->csr.d[0];
  A.gs = h->csr.d[1];
  A.max_ctas = h->prm.gn_max_ctas;
  A.slab_lo = h->slab_lo;
  A.slab_hi = h->slab_hi;
  // pointSearchInd is internal to the reference (LM:760, 867): the stage-level calls keep it for loam_map_get_corr, the
  // device loop does not write 20 bytes per stack point and iteration nobody reads
  A.nbr = nullptr;
  if (!solve) {
    LG_CHECK(h->gn.nbr.ensure((size_t)(A.n_cs + A.n_ss + 1) * 5 * 4, h->st));
    A.nbr = h->gn.nbr.as<int>();
  }
  h->nbr_valid = !solve;
  A.out = out_dev ? out_dev : h->d_mail;
  A.seq = out_dev ? 0ull : ++h->mail_seq;
  if (!out_dev) h->h_mail[40] = 0.0;
  if (use_px) {
    A.px = h->px;
    A.px.xseq = h->px.xseq + 1;  // sequence number of this launch's first iteration
  }
  return lg_map_gn_launch(h->gn, A, h->device, h->st, &h->launches);
}

// One iteration body without the solve (LM:754-967), sums through the mailbox.
int map_iter(loam_handle* h, int iter, const float* T, float* AtA, float* AtB, int* n_sel, bool use_px = false) {
  int rc = map_gn_enqueue(h, T, iter, iter + 1, 0, nullptr, use_px);
  if (rc) return rc;
  rc = mailbox_wait(h);
  if (rc) return rc;
  if (use_px) h->px.xseq++;
  h->d2h_bytes += 28 * 8;
  lg_unpack28(h->h_mail, AtA, AtB, n_sel);
  return LOAM_OK;
}

// LM:753-1017: the whole Gauss-Newton loop of one mapping run.  Default: on the device in one launch; the host then checks
// the eigen-decomposition of iteration 0 (LM:970-997) on the sums the kernel published, and only if that says "degenerate"
// (the kernel went on as if it did not) replays the run through the host, iteration by iteration.  LOAM_HOST_GN_LOOP=1
// forces the host path (bit-identical cross-check).
int map_optimize(loam_handle* h, float* Tt, int max_iters, int* iterations, bool use_px) {
  static const bool host_loop_env = getenv("LOAM_HOST_GN_LOOP") != nullptr;
  *iterations = 0;
  if (max_iters <= 0) return LOAM_OK;
  if (!host_loop_env) {
    int rc = map_gn_enqueue(h, Tt, 0, max_iters, 1, nullptr, use_px);
    if (rc) return rc;
    rc = mailbox_wait(h);
    if (rc) return rc;
    h->d2h_bytes += 41 * 8;
    const int last = (int)h->h_mail[38];
    if (use_px) h->px.xseq += (unsigned long long)(last + 1);
    bool speculation_held = true;
    if (h->h_mail[40] != 0.0) {  // iteration 0 had >= 50 rows: its eigen-decomposition decides matP / isDegenerate
      float AtA[36], AtB[6], X[6];
      int n_sel = 0;
      lg_unpack28(h->h_mail, AtA, AtB, &n_sel);
      const LgGNState saved = h->lm_gn;
      lg_gn_solve_step(AtA, AtB, 0, 100.f, h->lm_gn, X);
      if (h->lm_gn.degenerate) {
        h->lm_gn = saved;
        speculation_held = false;
      }
    }
    if (speculation_held) {
      for (int i = 0; i < 6; i++) Tt[i] = (float)h->h_mail[32 + i];
      *iterations = last + 1;
      return LOAM_OK;
    }
  }
  for (int iter = 0; iter < max_iters; iter++) {
    *iterations = iter + 1;
    float AtA[36], AtB[6], X[6];
    int n_sel = 0;
    int rc = map_iter(h, iter, Tt, AtA, AtB, &n_sel, use_px);
    if (rc) return rc;
    if (n_sel < 50) continue;  // LM:929-932
    lg_gn_solve_step(AtA, AtB, iter, 100.f, h->lm_gn, X);
    for (int i = 0; i < 6; i++) Tt[i] += X[i];
    float deltaR = (float)sqrt(pow(X[0] * 180.0 / M_PI, 2) + pow(X[1] * 180.0 / M_PI, 2) + pow(X[2] * 180.0 / M_PI, 2));
    float deltaT = (float)sqrt(pow(X[3] * 100, 2) + pow(X[4] * 100, 2) + pow(X[5] * 100, 2));
    if (deltaR < 0.05 && deltaT < 0.05) break;
  }
  return LOAM_OK;
}

// Box (in 1 m cells, inclusive) of the non-empty cubes among `cubes_idx`: a cube holds the points with
// int((p + 25) / 50) + cen == index (LM:1026-1032), i.e. floor(p) in [50 (index - cen) - 25, 50 (index - cen) + 24].
void cube_box(const loam_handle* h, const std::vector<std::vector<Chunk>>& cubes, const std::vector<int>& cubes_idx, int lo[3], int hi[3]) {
  for (int a = 0; a < 3; a++) lo[a] = INT32_MAX, hi[a] = INT32_MIN;
  for (int ind : cubes_idx) {
    long long n = 0;
    for (auto& c : cubes[ind]) n += c.n;
    if (!n) continue;
    const int c3[3] = {ind % CW - h->cenW, (ind / CW) % CH - h->cenH, ind / (CW * CH) - h->cenD};
    for (int a = 0; a < 3; a++) {
      lo[a] = std::min(lo[a], 50 * c3[a] - 25);
      hi[a] = std::max(hi[a], 50 * c3[a] + 24);
    }
  }
  if (lo[0] == INT32_MAX)
    for (int a = 0; a < 3; a++) lo[a] = hi[a] = 0;
}

ImuSC imu_sc(const float* v) {
  ImuSC s;
  s.s_pitch_s = sinf(v[0]); s.c_pitch_s = cosf(v[0]);
  s.s_yaw_s = sinf(v[1]); s.c_yaw_s = cosf(v[1]);
  s.s_roll_s = sinf(v[2]); s.c_roll_s = cosf(v[2]);
  s.s_pitch_l = sinf(v[3]); s.c_pitch_l = cosf(v[3]);
  s.s_yaw_l = sinf(v[4]); s.c_yaw_l = cosf(v[4]);
  s.s_roll_l = sinf(v[5]); s.c_roll_l = cosf(v[5]);
  s.shift[0] = v[6]; s.shift[1] = v[7]; s.shift[2] = v[8];
  return s;
}

}  // namespace

// ===================================================================================================== lifecycle
extern "C" {

const char* loam_strerror(int code) {
  switch (code) {
    case LOAM_OK: return "ok";
    case LOAM_EINVAL: return "invalid argument";
    case LOAM_ECUDA: return "CUDA error";
    case LOAM_ENOSPC: return "buffer or capacity too small";
    case LOAM_ESTATE: return "call order violated";
    case LOAM_EUNSUPPORTED: return "unsupported input";
  }
  return "unknown error";
}
const char* loam_last_cuda_error(const loam_handle*) { return g_cuda_err; }

void loam_default_params(loam_params* p) {
  p->n_scans = 16;
  p->ring_mode = 0;
  p->ring_ang_min = -15.f;
  p->ring_ang_step = 2.f;
  p->skip_frame_num = 1;
  p->max_points = 131072;
  p->max_map_points = 1 << 21;
  p->want_registered = 0;
  p->want_surround = 0;
  p->pose_message_hop = 0;
  p->gn_max_ctas = 0;
}

// role bits: 1 = needs the per-sweep odometry buffers, 2 = needs the map storage (arena, sort workspace)
static int create_internal(const loam_params* p, int device, int role, loam_handle** out) {
  if (!out) return LOAM_EINVAL;
  *out = nullptr;
  loam_params prm;
  if (p) prm = *p; else loam_default_params(&prm);
  if (prm.n_scans < 1 || prm.n_scans > 64) return LOAM_EINVAL;
  int ndev = 0;
  LG_CHECK(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return LOAM_EINVAL;
  LG_CHECK(cudaSetDevice(device));
  loam_handle* h = new loam_handle;
  h->prm = prm;
  h->device = device;
  h->srp.n_scans = prm.n_scans;
  h->srp.ring_mode = prm.ring_mode;
  h->srp.ring_ang_min = prm.ring_ang_min;
  h->srp.ring_ang_step = prm.ring_ang_step;
  h->srp.scan_period = 0.1;
  h->frameCount = prm.skip_frame_num;
  h->cubeC.resize(CNUM);
  h->cubeS.resize(CNUM);
  cudaError_t e = cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaHostAlloc((void**)&h->h_mail, 64 * sizeof(double), cudaHostAllocMapped);
  if (e == cudaSuccess) e = cudaHostGetDevicePointer((void**)&h->d_mail, h->h_mail, 0);
  if (e == cudaSuccess) e = cudaHostAlloc((void**)&h->h_ints, loam_handle::H_INTS * sizeof(int), cudaHostAllocDefault);
  if (e == cudaSuccess) e = cudaHostAlloc((void**)&h->h_ints2, loam_handle::H_INTS * sizeof(int), cudaHostAllocDefault);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->pds.ev, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_pre_ds, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->st2, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    lg_set_error(cudaGetErrorString(e), __FILE__, __LINE__);
    delete h;
    return LOAM_ECUDA;
  }
  memset(h->h_mail, 0, 64 * sizeof(double));  // the sequence word must not start at a recycled page's stale value
  // reserve the steady-state working set up front (HBM is plentiful; reallocation stalls are not)
  {
    const size_t mp = (size_t)std::max(prm.max_points, 1024), mm = (size_t)std::max(prm.max_map_points, 1024);
    DevBuf* sweep_bufs[] = {&h->corner_last, &h->surf_last, &h->corner_new, &h->surf_new, &h->fullres3, &h->stack2_c, &h->stack2_s,
                            &h->stack_c, &h->stack_s, &h->ins_sel, &h->ins_sorted};
    if (role & 1)
      for (DevBuf* b : sweep_bufs) e = e == cudaSuccess ? b->ensure(mp * 16, h->st) : e;
    DevBuf* map_bufs[] = {&h->map_c, &h->map_s, &h->ds_in};
    if (role & 2) {
      for (DevBuf* b : map_bufs) e = e == cudaSuccess ? b->ensure(mm * 16, h->st) : e;
      if (e == cudaSuccess) e = h->arena.ensure(mm * 4 * 16, h->st);
      if (e == cudaSuccess) e = h->arena2.ensure(mm * 4 * 16, h->st);
      if (e == cudaSuccess) e = (cudaError_t)lg_radix_ensure(h->vb.rs, (int)mm, h->st) == cudaSuccess ? cudaSuccess : cudaErrorMemoryAllocation;
      // cell-sorted index of the local map (tables for a 250 x 100 x 250 m box twice) and the merge-path key arrays
      if (e == cudaSuccess && lg_csr_reserve(h->csr, (size_t)16 << 20, (int)mm, h->st) != LOAM_OK) e = cudaErrorMemoryAllocation;
      DevBuf* key_bufs[] = {&h->vb.keys_old, &h->vb.keys_m};
      DevBuf* val_bufs[] = {&h->vb.vals_old, &h->vb.vals_m};
      for (DevBuf* b : key_bufs) e = e == cudaSuccess ? b->ensure(mm * 8, h->st) : e;
      for (DevBuf* b : val_bufs) e = e == cudaSuccess ? b->ensure(mm * 4, h->st) : e;
    }
    if (role & 4) {  // output handle of a pipeline: surround cloud scratch (gathered cubes, sort workspace, result)
      DevBuf* sur_bufs[] = {&h->vg_in, &h->ds_in, &h->vg_out, &h->surround};
      for (DevBuf* b : sur_bufs) e = e == cudaSuccess ? b->ensure(mm * 16, h->st) : e;
      if (e == cudaSuccess) e = (cudaError_t)lg_radix_ensure(h->vb.rs, (int)mm, h->st) == cudaSuccess ? cudaSuccess : cudaErrorMemoryAllocation;
      if (e == cudaSuccess) e = h->d_ents.ensure(4096 * sizeof(CopyEnt), h->st);
    }
    if (e != cudaSuccess) {
      lg_set_error(cudaGetErrorString(e), __FILE__, __LINE__);
      loam_destroy(h);
      return LOAM_ECUDA;
    }
  }
  *out = h;
  return LOAM_OK;
}

int loam_create(const loam_params* p, int device, loam_handle** out) { return create_internal(p, device, 3, out); }

int loam_destroy(loam_handle* h) {
  if (!h) return LOAM_EINVAL;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->st);
  if (h->st2) cudaStreamSynchronize(h->st2);
  h->prof.resolve(h->st);
  h->prof.release();
  h->sr.release(); h->od.release(); h->csr.release(); h->gn.release(); h->vb.release();
  DevBuf* all[] = {&h->xyz_packed, &h->wire, &h->xyz_in, &h->t_sharp, &h->t_flat, &h->corner_last, &h->surf_last, &h->corner_new, &h->surf_new, &h->fullres3,
                   &h->arena, &h->arena2, &h->stack2_c, &h->stack2_s, &h->stack_c, &h->stack_s, &h->map_c, &h->map_s, &h->d_ents, &h->d_segs,
                   &h->d_ints, &h->d_seg_off, &h->d_seg_leaf, &h->d_out_se, &h->ds_in, &h->ins_sel, &h->ins_sorted, &h->d_runs,
                   &h->surround, &h->registered, &h->vg_in, &h->vg_out, &h->vs_staging, &h->vs_counts, &h->batch_tab, &h->d_imu_ring, &h->d_imu_carry};
  for (DevBuf* b : all) b->release();
  if (h->px_connected)
    for (int r = 0; r < h->px.world; r++)
      if (r != h->px.rank && h->px.buf[r] && !h->px_local[r]) cudaIpcCloseMemHandle(h->px.buf[r]);
  if (h->xchg) {
    std::lock_guard<std::mutex> lock(g_xchg_mutex);
    for (auto it = g_xchg_local.begin(); it != g_xchg_local.end();) it = (it->second == h->xchg) ? g_xchg_local.erase(it) : std::next(it);
    cudaFree(h->xchg);
  }
  if (h->ev_map_done) cudaEventDestroy(h->ev_map_done);
  if (h->ev_aux_read) cudaEventDestroy(h->ev_aux_read);
  if (h->h_mail) cudaFreeHost(h->h_mail);
  if (h->h_ints) cudaFreeHost(h->h_ints);
  if (h->h_ints2) cudaFreeHost(h->h_ints2);
  if (h->pds.ev) cudaEventDestroy(h->pds.ev);
  if (h->ev_pre_ds) cudaEventDestroy(h->ev_pre_ds);
  if (h->st2) cudaStreamDestroy(h->st2);
  if (h->st) cudaStreamDestroy(h->st);
  delete h;
  return LOAM_OK;
}

int loam_reset(loam_handle* h) {
  if (!h) return LOAM_EINVAL;
  h->lo_inited = false;  // LO:411-415; everything else follows on the next sweeps
  return LOAM_OK;
}
void* loam_stream(loam_handle* h) { return h ? (void*)h->st : nullptr; }
int loam_stats(const loam_handle* h, long long out4[4]) {
  if (!h || !out4) return LOAM_EINVAL;
  out4[0] = h->launches;
  out4[1] = h->h2d_bytes;
  out4[2] = h->d2h_bytes;
  out4[3] = h->syncs;
  return LOAM_OK;
}
int loam_host_times(loam_handle* h, double* out16, int clear) {
  if (!h || !out16) return LOAM_EINVAL;
  for (int i = 0; i < LOAM_HOST_SECTIONS; i++) {
    out16[i] = h->host_s[i];
    if (clear) h->host_s[i] = 0.0;
  }
  return LOAM_OK;
}
int loam_launch_latency(loam_handle* h, int n, double* period_us, double* roundtrip_us) {
  if (!h || n < 1 || !period_us || !roundtrip_us) return LOAM_EINVAL;
  LG_CHECK(cudaSetDevice(h->device));
  cudaEvent_t a, b;
  LG_CHECK(cudaEventCreate(&a));
  LG_CHECK(cudaEventCreate(&b));
  for (int i = 0; i < 16; i++) lg_empty_launch(h->st);
  LG_CHECK(cudaStreamSynchronize(h->st));
  LG_CHECK(cudaEventRecord(a, h->st));
  for (int i = 0; i < n; i++) lg_empty_launch(h->st);
  LG_CHECK(cudaEventRecord(b, h->st));
  LG_CHECK(cudaStreamSynchronize(h->st));
  float ms = 0.f;
  LG_CHECK(cudaEventElapsedTime(&ms, a, b));
  *period_us = 1e3 * ms / n;
  const int m = n < 200 ? n : 200;
  const auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < m; i++) {
    lg_empty_launch(h->st);
    LG_CHECK(cudaStreamSynchronize(h->st));
  }
  *roundtrip_us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / m;
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  return LOAM_OK;
}

int loam_profile(loam_handle* h, int enable) {
  if (!h) return LOAM_EINVAL;
  cudaSetDevice(h->device);
  h->prof.resolve(h->st);
  h->prof.on = enable != 0;
  for (int i = 0; i < LGK_COUNT; i++) h->prof.ms[i] = h->prof.units[i] = 0.0, h->prof.scopes[i] = 0;
  return LOAM_OK;
}
int loam_profile_read(loam_handle* h, double* ms, double* units, long long* scopes, int n) {
  if (!h || n < LGK_COUNT) return LOAM_EINVAL;
  cudaSetDevice(h->device);
  h->prof.resolve(h->st);
  for (int i = 0; i < LGK_COUNT; i++) {
    if (ms) ms[i] = h->prof.ms[i];
    if (units) units[i] = h->prof.units[i];
    if (scopes) scopes[i] = h->prof.scopes[i];
  }
  return LOAM_OK;
}
long long loam_launch_count(const loam_handle* h) { return h ? h->launches : 0; }

// ============================================================================================ scanRegistration
int loam_extract(loam_handle* h, const float* xyz_host, int n, int stride_bytes, double stamp, const float* imu_trans, loam_counts* out) {
  if (!h || (!xyz_host && n > 0)) return LOAM_EINVAL;
  LG_CHECK(cudaSetDevice(h->device));
  g_lg_prof = h->prof.on ? &h->prof : nullptr;
  int rc = upload(h, h->xyz_in, xyz_host, (size_t)n * stride_bytes);
  if (rc) return rc;
  return extract_common(h, h->xyz_in.as<float>(), n, stride_bytes, imu_trans, out, stamp);
}
int loam_extract_device(loam_handle* h, const float* xyz_dev, int n, int stride_bytes, double stamp, const float* imu_trans, loam_counts* out) {
  if (!h || (!xyz_dev && n > 0)) return LOAM_EINVAL;
  LG_CHECK(cudaSetDevice(h->device));
  g_lg_prof = h->prof.on ? &h->prof : nullptr;
  return extract_common(h, xyz_dev, n, stride_bytes, imu_trans, out, stamp);
}

// imuHandler SR:754-837: one /imu/data message.  From the first message on, loam_extract de-skews every sweep (SR:364-434)
// and produces /imu_trans itself.
int loam_imu_push(loam_handle* h, double stamp, const double* orientation_xyzw, const double* angular_velocity, const double* linear_acceleration) {
  if (!h || !orientation_xyzw || !angular_velocity || !linear_acceleration) return LOAM_EINVAL;
  lgh::imu_handler(h->imu_host, h->imu_ring, stamp, orientation_xyzw, angular_velocity, linear_acceleration);
  h->imu_ring_dirty = true;
  return LOAM_OK;
}
int loam_get_imu_trans(loam_handle* h, float* out12) {
  if (!h || !out12) return LOAM_EINVAL;
  for (int i = 0; i < 12; i++) out12[i] = h->imu[i];
  return LOAM_OK;
}

// One sweep of each of B independent sequences (SURVEY 8b `*_batch`): the eight extraction kernels are launched ONCE for
// all members (grid.y = sequence) on the first handle's stream, the counts come back with one synchronisation.  Results
// are those of B loam_extract calls, bit for bit.  Members the batched launch does not take (empty sweeps, unaligned
// point layouts) and sweeps with virtual rings go through the per-handle path.
int loam_extract_batch(loam_handle* const* hs, int B, const float* const* xyz_host, const int* n, int stride_bytes, const double* stamps,
                       loam_counts* out) {
  if (!hs || !xyz_host || !n || B < 1 || B > 256 || stride_bytes < 12) return LOAM_EINVAL;
  for (int b = 0; b < B; b++)
    if (!hs[b] || n[b] < 0 || (!xyz_host[b] && n[b] > 0) || hs[b]->device != hs[0]->device || hs[b]->prm.n_scans != hs[0]->prm.n_scans)
      return LOAM_EINVAL;
  for (int b = 0; b < B; b++)
    for (int c = 0; c < b; c++)
      if (hs[b] == hs[c]) return LOAM_EINVAL;
  loam_handle* h0 = hs[0];
  LG_CHECK(cudaSetDevice(h0->device));
  g_lg_prof = h0->prof.on ? &h0->prof : nullptr;
  HostTimer ht(&h0->host_s[HT_EXTRACT]);
  cudaStream_t st = h0->st;
  std::vector<int> members;
  std::vector<SrWs*> ws;
  std::vector<SrParams> prm;
  std::vector<const float*> dxyz;
  std::vector<int> nn, strides;
  for (int b = 0; b < B; b++) {
    loam_handle* h = hs[b];
    const bool batched = n[b] > 0 && (stride_bytes & 3) == 0 && h->imu_host.last < 0;  // the IMU de-skew is a per-handle pass
    if (!batched) {  // per-handle path
      int rc = loam_extract(h, xyz_host[b], n[b], stride_bytes, stamps ? stamps[b] : 0.0, nullptr, out ? &out[b] : nullptr);
      if (rc) return rc;
      continue;
    }
    LG_CHECK(cudaStreamSynchronize(h->st));  // the member's own stream may still read the clouds of its previous sweep
    int rc = upload_on(h, st, h->xyz_in, xyz_host[b], (size_t)n[b] * stride_bytes);
    if (rc) return rc;
    for (int i = 0; i < 12; i++) h->imu[i] = 0.f;
    members.push_back(b);
    ws.push_back(&h->sr);
    prm.push_back(h->srp);
    dxyz.push_back(h->xyz_in.as<float>());
    nn.push_back(n[b]);
    strides.push_back(stride_bytes);
  }
  if (members.empty()) return LOAM_OK;
  int rc = lg_extract_launch_batch(ws.data(), prm.data(), dxyz.data(), nn.data(), strides.data(), (int)members.size(), h0->batch_tab, st,
                                   &h0->launches);
  if (rc) return rc;
  for (int b : members) {
    loam_handle* h = hs[b];
    h->d2h_bytes += SRM_HEAD * 4;
    LG_CHECK(cudaMemcpyAsync(h->h_ints, h->sr.meta.p, SRM_HEAD * 4, cudaMemcpyDeviceToHost, st));
  }
  h0->syncs++;
  LG_CHECK(cudaStreamSynchronize(st));
  for (int b : members) {
    loam_handle* h = hs[b];
    g_lg_prof = h->prof.on ? &h->prof : nullptr;
    rc = sr_counts_tail(h, out ? &out[b] : nullptr);  // sweeps with virtual rings are finished on the member's own stream
    if (rc) return rc;
    extract_publish(h);
  }
  return LOAM_OK;
}

// ============================================================================================ laserOdometry
// LO:519-572 -- the part of a sweep before its Gauss-Newton iterations.  *state: 0 = the sweep only initialised the node
// (LO:519-563), 1 = iterate (LO:572 holds), 2 = go straight to the pose accumulation.  Device work goes to `st`.
static int lo_begin(loam_handle* h, loam_odom_result* out, cudaStream_t st, int* state) {
  memset(out, 0, sizeof(*out));
  const loam_counts& c = h->counts;
  const float* imu = h->imu;
  if (!h->lo_inited) {  // LO:519-563
    h->cornerLastNum = 0;
    h->surfLastNum = 0;
    LG_CHECK(h->corner_last.ensure((size_t)(c.n_less_sharp + 16) * 16, st));
    LG_CHECK(h->surf_last.ensure((size_t)(c.n_less_flat + 16) * 16, st));
    if (c.n_less_sharp) LG_CHECK(cudaMemcpyAsync(h->corner_last.p, h->cur_less_sharp, (size_t)c.n_less_sharp * 16, cudaMemcpyDeviceToDevice, st));
    if (c.n_less_flat) LG_CHECK(cudaMemcpyAsync(h->surf_last.p, h->cur_less_flat, (size_t)c.n_less_flat * 16, cudaMemcpyDeviceToDevice, st));
    h->n_corner_last = c.n_less_sharp;
    h->n_surf_last = c.n_less_flat;
    h->od.bounds_valid = false;
    for (int i = 0; i < 6; i++) h->T[i] = h->Tsum[i] = 0.f;
    h->Tsum[0] += imu[0];  // imuPitchStart
    h->Tsum[2] += imu[2];  // imuRollStart
    h->lo_inited = true;
    out->clouds_published = 1;
    out->n_corner_last = h->n_corner_last;
    out->n_surf_last = h->n_surf_last;
    *state = 0;
    return LOAM_OK;
  }
  const float scanPeriod = 0.1f;  // LO:50
  h->T[3] -= imu[9] * scanPeriod;
  h->T[4] -= imu[10] * scanPeriod;
  h->T[5] -= imu[11] * scanPeriod;
  *state = (h->cornerLastNum > 10 && h->surfLastNum > 100) ? 1 : 2;  // LO:572
  return LOAM_OK;
}

// LO:904-907, 975-1031 for iteration 0 on the host (it carries the eigen-decomposition / degeneracy test): true = converged
static bool lo_iter0_host(loam_handle* h, const float* AtA, const float* AtB, int n_sel) {
  if (n_sel < 10) return false;  // LO:904-907
  float X[6];
  lg_gn_solve_step(AtA, AtB, 0, 10.f, h->lo_gn, X);
  for (int i = 0; i < 6; i++) h->T[i] += X[i];
  for (int i = 0; i < 6; i++)
    if (isnan(h->T[i])) h->T[i] = 0;
  float deltaR = (float)sqrt(pow(X[0] * 180.0 / M_PI, 2) + pow(X[1] * 180.0 / M_PI, 2) + pow(X[2] * 180.0 / M_PI, 2));
  float deltaT = (float)sqrt(pow(X[3] * 100, 2) + pow(X[4] * 100, 2) + pow(X[5] * 100, 2));
  return deltaR < 0.1 && deltaT < 0.1;
}

// LO:1035-1121 -- pose accumulation, TransformToEnd of the clouds, swap.  k == nullptr: launches the kernel on `st`; else the
// kernel arguments go into the member's row of a batched launch (lg_odom_batch_to_end on the same stream).
static int lo_finish(loam_handle* h, loam_odom_result* out, cudaStream_t st, OdK* k) {
  const loam_counts& c = h->counts;
  const float* imu = h->imu;
  // LO:1035-1064
  float* T = h->T;
  float* S = h->Tsum;
  float rx, ry, rz, tx, ty, tz;
  lgh::accumulate_rotation(S[0], S[1], S[2], -T[0], (float)(-T[1] * 1.05), -T[2], rx, ry, rz);
  float x1 = cosf(rz) * (T[3] - imu[6]) - sinf(rz) * (T[4] - imu[7]);
  float y1 = sinf(rz) * (T[3] - imu[6]) + cosf(rz) * (T[4] - imu[7]);
  float z1 = (float)(T[5] * 1.05 - imu[8]);
  float x2 = x1;
  float y2 = cosf(rx) * y1 - sinf(rx) * z1;
  float z2 = sinf(rx) * y1 + cosf(rx) * z1;
  tx = S[3] - (cosf(ry) * x2 + sinf(ry) * z2);
  ty = S[4] - y2;
  tz = S[5] - (-sinf(ry) * x2 + cosf(ry) * z2);
  lgh::plugin_imu_rotation(rx, ry, rz, imu[0], imu[1], imu[2], imu[3], imu[4], imu[5], rx, ry, rz);
  S[0] = rx; S[1] = ry; S[2] = rz; S[3] = tx; S[4] = ty; S[5] = tz;
  out->odom_published = 1;

  // LO:1087-1121
  h->frameCount++;
  const bool pub = h->frameCount >= h->prm.skip_frame_num + 1;
  LG_CHECK(h->corner_new.ensure((size_t)(c.n_less_sharp + 16) * 16, st));
  LG_CHECK(h->surf_new.ensure((size_t)(c.n_less_flat + 16) * 16, st));
  if (pub) LG_CHECK(h->fullres3.ensure((size_t)(c.n_full + 16) * 16, st));
  OdomT ot;
  for (int i = 0; i < 6; i++) ot.t[i] = T[i];
  if (k) {
    k->do_to_end = 1;
    k->T = ot; k->sT = host_sincos3(T); k->imu = imu_sc(imu);
    k->in0 = h->cur_less_sharp; k->out0 = h->corner_new.as<float4>(); k->n0 = c.n_less_sharp;
    k->in1 = h->cur_less_flat; k->out1 = h->surf_new.as<float4>(); k->n1 = c.n_less_flat;
    k->in2 = h->cur_full; k->out2 = h->fullres3.as<float4>(); k->n2 = pub ? c.n_full : 0;
  } else {
    int rc = lg_odom_to_end_launch(ot, host_sincos3(T), imu_sc(imu), h->cur_less_sharp, h->corner_new.as<float4>(), c.n_less_sharp,
                                   h->cur_less_flat, h->surf_new.as<float4>(), c.n_less_flat, h->cur_full, h->fullres3.as<float4>(),
                                   pub ? c.n_full : 0, st, &h->launches);
    if (rc) return rc;
  }
  std::swap(h->corner_last, h->corner_new);
  std::swap(h->surf_last, h->surf_new);
  h->od.bounds_valid = false;
  h->n_corner_last = c.n_less_sharp;
  h->n_surf_last = c.n_less_flat;
  h->cornerLastNum = h->n_corner_last;
  h->surfLastNum = h->n_surf_last;
  if (pub) {
    h->frameCount = 0;
    h->n_fullres3 = c.n_full;
    out->clouds_published = 1;
    out->fullres_published = 1;
  }
  for (int i = 0; i < 6; i++) {
    out->transform_sum[i] = S[i];
    out->transformation[i] = T[i];
  }
  out->n_corner_last = h->n_corner_last;
  out->n_surf_last = h->n_surf_last;
  return LOAM_OK;
}

int loam_odometry_process(loam_handle* h, loam_odom_result* out) {
  if (!h || !out) return LOAM_EINVAL;
  if (!h->have_features) return LOAM_ESTATE;
  LG_CHECK(cudaSetDevice(h->device));
  g_lg_prof = h->prof.on ? &h->prof : nullptr;
  int state = 0;
  {
    int rcb = lo_begin(h, out, h->st, &state);
    if (rcb) return rcb;
    if (state == 0) return LOAM_OK;
  }
  const loam_counts& c = h->counts;
  HostTimer ht(&h->host_s[HT_ODOM_ITERS]);
  if (state == 1) {  // LO:572
    static const bool host_loop = getenv("LOAM_HOST_GN_LOOP") != nullptr;  // diagnostic: every iteration through the host
    for (int iter = 0; iter < 25;) {
      if (iter == 0 || host_loop) {  // iteration 0 carries the eigen-decomposition / degeneracy test (LO:977-999): host
        out->iterations = iter + 1;
        float AtA[36], AtB[6], X[6];
        int n_sel = 0;
        int rc = odom_iter(h, iter, h->T, AtA, AtB, &n_sel);
        if (rc) return rc;
        iter++;
        if (!host_loop || iter == 1) {
          if (lo_iter0_host(h, AtA, AtB, n_sel)) break;
          continue;
        }
        if (n_sel < 10) continue;  // LO:904-907 (diagnostic path: later iterations through the host)
        lg_gn_solve_step(AtA, AtB, iter - 1, 10.f, h->lo_gn, X);
        for (int i = 0; i < 6; i++) h->T[i] += X[i];
        for (int i = 0; i < 6; i++)
          if (isnan(h->T[i])) h->T[i] = 0;
        float deltaR = (float)sqrt(pow(X[0] * 180.0 / M_PI, 2) + pow(X[1] * 180.0 / M_PI, 2) + pow(X[2] * 180.0 / M_PI, 2));
        float deltaT = (float)sqrt(pow(X[3] * 100, 2) + pow(X[4] * 100, 2) + pow(X[5] * 100, 2));
        if (deltaR < 0.1 && deltaT < 0.1) break;
        continue;
      }
      // iterations up to the next correspondence refresh run on the device without the host in between
      OdomLoopArgs la;
      for (int i = 0; i < 6; i++) la.T.t[i] = h->T[i];
      la.sc = host_sincos3(h->T);
      memcpy(la.matP, h->lo_gn.matP, sizeof(la.matP));
      la.degenerate = h->lo_gn.degenerate ? 1 : 0;
      la.it0 = iter;
      la.it1 = std::min(25, (iter / 5 + 1) * 5);
      int rc = lg_odom_loop_launch(h->od, la, h->cur_sharp, c.n_sharp, h->cur_flat, c.n_flat, h->corner_last.as<float4>(), h->n_corner_last,
                                   h->surf_last.as<float4>(), h->n_surf_last, h->d_mail, ++h->mail_seq, h->st, &h->launches);
      if (rc) return rc;
      rc = mailbox_wait(h);
      if (rc) return rc;
      h->d2h_bytes += 8 * 8;
      for (int i = 0; i < 6; i++) h->T[i] = (float)h->h_mail[i];
      out->iterations = (int)h->h_mail[6] + 1;
      if (h->h_mail[7] != 0.0) break;
      iter = la.it1;
    }
  }
  ht.lap(&h->host_s[HT_ODOM_END]);
  return lo_finish(h, out, h->st, nullptr);
}

// SURVEY 8b `*_batch`: one sweep of B independent sequences through scan-to-scan odometry in LOCK-STEP.  Every kernel of a
// round (box hierarchy, correspondence refresh, iteration 0, a block of device iterations, TransformToEnd) is launched once
// for all members that take part in it (grid.y = sequence); the host parts (LO:519-572, the eigen-decomposition of
// iteration 0, LO:1035-1084) run per member between the rounds.  Results are those of B loam_odometry_process calls, bit
// for bit.  Members whose feature count exceeds what the cluster kernels hold go through the per-handle call.
int loam_odometry_process_batch(loam_handle* const* hs, int B, loam_odom_result* out) {
  if (!hs || !out || B < 1 || B > 256) return LOAM_EINVAL;
  for (int b = 0; b < B; b++) {
    if (!hs[b] || hs[b]->device != hs[0]->device) return LOAM_EINVAL;
    if (!hs[b]->have_features) return LOAM_ESTATE;
    for (int c = 0; c < b; c++)
      if (hs[b] == hs[c]) return LOAM_EINVAL;
  }
  loam_handle* h0 = hs[0];
  LG_CHECK(cudaSetDevice(h0->device));
  g_lg_prof = h0->prof.on ? &h0->prof : nullptr;
  HostTimer ht(&h0->host_s[HT_ODOM_ITERS]);
  cudaStream_t st = h0->st;
  std::vector<OdK> tab(B);
  std::vector<int> state(B, 0), iter(B, 0);
  std::vector<char> active(B, 0), batched(B, 0);
  for (int b = 0; b < B; b++) {
    loam_handle* h = hs[b];
    memset(&tab[b], 0, sizeof(OdK));
    if (!lg_odom_batch_fits(h->counts.n_sharp, h->counts.n_flat)) {  // per-handle path (own stream, blocking)
      int rc = loam_odometry_process(h, &out[b]);
      if (rc) return rc;
      continue;
    }
    batched[b] = 1;
    if (h != h0) LG_CHECK(cudaStreamSynchronize(h->st));  // the member's own stream may still use the clouds this call replaces
    int rc = lo_begin(h, &out[b], st, &state[b]);
    if (rc) return rc;
    if (state[b] != 1) continue;
    OdK& k = tab[b];
    rc = lg_odom_batch_prepare(h->od, h->counts.n_sharp, h->counts.n_flat, h->n_corner_last, h->n_surf_last, st, &k);
    if (rc) return rc;
    k.sharp = h->cur_sharp; k.flat = h->cur_flat; k.corner_last = h->corner_last.as<float4>(); k.surf_last = h->surf_last.as<float4>();
    k.n_sharp = h->counts.n_sharp; k.n_flat = h->counts.n_flat; k.n_cl = h->n_corner_last; k.n_sl = h->n_surf_last;
    k.out = h->d_mail;
    active[b] = 1;
  }
  // ---- round 0: box hierarchy (once per sweep), correspondences, iteration 0 (its sums go to the host: LO:977-999)
  bool any = false;
  for (int b = 0; b < B; b++) {
    OdK& k = tab[b];
    k.do_bounds = k.do_refresh = k.do_iter0 = k.do_loop = 0;
    if (!active[b]) continue;
    loam_handle* h = hs[b];
    any = true;
    k.do_bounds = h->od.bounds_valid ? 0 : 1;
    k.do_refresh = 1;
    k.do_iter0 = 1;
    for (int i = 0; i < 6; i++) k.T.t[i] = h->T[i];
    k.sc = host_sincos3(h->T);
    k.iter = 0;
    k.seq = ++h->mail_seq;
  }
  if (any) {
    int rc = lg_odom_batch_round(tab.data(), B, h0->batch_tab, st, &h0->launches);
    if (rc) return rc;
    for (int b = 0; b < B; b++) {
      if (!active[b]) continue;
      loam_handle* h = hs[b];
      h->od.bounds_valid = true;
      rc = mailbox_wait(h, st);
      if (rc) return rc;
      h->d2h_bytes += 28 * 8;
      float AtA[36], AtB[6];
      int n_sel = 0;
      lg_unpack28(h->h_mail, AtA, AtB, &n_sel);
      out[b].iterations = 1;
      iter[b] = 1;
      if (lo_iter0_host(h, AtA, AtB, n_sel)) active[b] = 0;
    }
  }
  // ---- rounds 1..: blocks of device iterations up to the next correspondence refresh (LO:595), members drop out as they converge
  for (;;) {
    any = false;
    for (int b = 0; b < B; b++) {
      OdK& k = tab[b];
      k.do_bounds = k.do_refresh = k.do_iter0 = k.do_loop = 0;
      if (!active[b]) continue;
      loam_handle* h = hs[b];
      any = true;
      OdomLoopArgs& la = k.la;
      for (int i = 0; i < 6; i++) la.T.t[i] = h->T[i];
      la.sc = host_sincos3(h->T);
      memcpy(la.matP, h->lo_gn.matP, sizeof(la.matP));
      la.degenerate = h->lo_gn.degenerate ? 1 : 0;
      la.it0 = iter[b];
      la.it1 = std::min(25, (iter[b] / 5 + 1) * 5);
      k.T = la.T;
      k.do_refresh = la.it0 % 5 == 0 ? 1 : 0;
      k.do_loop = 1;
      k.seq = ++h->mail_seq;
    }
    if (!any) break;
    int rc = lg_odom_batch_round(tab.data(), B, h0->batch_tab, st, &h0->launches);
    if (rc) return rc;
    for (int b = 0; b < B; b++) {
      if (!active[b]) continue;
      loam_handle* h = hs[b];
      rc = mailbox_wait(h, st);
      if (rc) return rc;
      h->d2h_bytes += 8 * 8;
      for (int i = 0; i < 6; i++) h->T[i] = (float)h->h_mail[i];
      out[b].iterations = (int)h->h_mail[6] + 1;
      iter[b] = tab[b].la.it1;
      if (h->h_mail[7] != 0.0 || iter[b] >= 25) active[b] = 0;
    }
  }
  ht.lap(&h0->host_s[HT_ODOM_END]);
  // ---- LO:1035-1121 per member on the host, TransformToEnd of all members in one launch
  for (int b = 0; b < B; b++) {
    tab[b].do_to_end = 0;
    if (!batched[b] || state[b] == 0) continue;
    int rc = lo_finish(hs[b], &out[b], st, &tab[b]);
    if (rc) return rc;
  }
  int rc = lg_odom_batch_to_end(tab.data(), B, h0->batch_tab, st, &h0->launches);
  if (rc) return rc;
  h0->syncs++;
  LG_CHECK(cudaStreamSynchronize(st));  // the members go on with their own streams
  return LOAM_OK;
}

// ============================================================================================ laserMapping
int loam_pose_message_hop(const float* in6, float* out6) {
  if (!in6 || !out6) return LOAM_EINVAL;
  lgh::pose_message_hop(in6, out6);
  return LOAM_OK;
}

int loam_mapping_odometry(loam_handle* h, const float* Tsum) {
  if (!h || !Tsum) return LOAM_EINVAL;
  if (fabs((double)Tsum[3]) < 0.000001 && fabs((double)Tsum[4]) < 0.000001 && fabs((double)Tsum[5]) < 0.000001) h->lm_inited = false;  // LM:316-319
  for (int i = 0; i < 6; i++) h->mTsum[i] = Tsum[i];
  return LOAM_OK;
}

// Applies the read-back of the previous run's per-cube voxel grids (see loam_handle::PendingDS).
static int finish_cube_ds(loam_handle* h) {
  loam_handle::PendingDS& pd = h->pds;
  if (!pd.active) return LOAM_OK;
  pd.active = false;
  h->syncs++;
  LG_CHECK(cudaEventSynchronize(pd.ev));
  const int nseg = pd.nseg;
  if (pd.merged && h->h_ints2[2 * nseg] != 0) {  // a cube's old cloud was not in voxel order after all: full sort path
    h->merge_fallbacks++;
    int* d_start = h->d_out_se.as<int>();
    int* d_end = d_start + nseg;
    int rc = upload(h, h->d_ents, pd.ents.data(), pd.ents.size() * sizeof(CopyEnt));
    if (rc) return rc;
    rc = lg_gather(h->d_ents.as<CopyEnt>(), (int)pd.ents.size(), pd.max_n, h->ds_in.as<float4>(), h->st, &h->launches);
    if (rc) return rc;
    rc = upload(h, h->d_seg_off, pd.seg_off.data(), (nseg + 1) * 4);
    if (rc) return rc;
    rc = upload(h, h->d_seg_leaf, pd.leaf.data(), nseg * 4);
    if (rc) return rc;
    LG_CHECK(h->ds_in.ensure((size_t)(pd.Mtot + 16) * 16, h->st));
    rc = lg_vox_big(h->vb, h->ds_in.as<float4>(), h->d_seg_off.as<int>(), h->d_seg_leaf.as<float>(), nseg, pd.Mtot,
                    h->arena.as<float4>() + h->bump, d_start, d_end, h->st, &h->launches);
    if (rc) return rc;
    LG_D2H(h, h->h_ints2, d_start, (size_t)nseg * 8);
    LG_SYNC(h);
  }
  int total = 0, s = 0;
  for (int ind : pd.validInd)
    for (int type = 0; type < 2; type++) {
      auto& cube = type == 0 ? h->cubeC[ind] : h->cubeS[ind];
      int st0 = h->h_ints2[s], en0 = h->h_ints2[nseg + s];
      cube.clear();
      if (en0 > st0) cube.push_back(Chunk{(int)h->bump + st0, en0 - st0});
      total = std::max(total, en0);
      s++;
    }
  h->bump += total;
  return LOAM_OK;
}

int loam_mapping_process(loam_handle* h, loam_map_result* out) {
  if (!h || !out) return LOAM_EINVAL;
  LG_CHECK(cudaSetDevice(h->device));
  g_lg_prof = h->prof.on ? &h->prof : nullptr;
  memset(out, 0, sizeof(*out));
  HostTimer ht(&h->host_s[HT_MAP_PREP]);
  if (!h->lm_inited) {  // a reset clears the cubes: bring them up to date first
    int rcf = finish_cube_ds(h);
    if (rcf) return rcf;
  }
  if (h->aux_read_pending) {  // the output stream may still be reading cubes of the previous surround: order this run's arena writes behind it
    LG_CHECK(cudaStreamWaitEvent(h->st, h->ev_aux_read, 0));
    h->aux_read_pending = false;
  }
  if (!h->lm_inited) {
    h->lm_inited = true;
    map_reset(h);
  }
  lgh::transform_associate_to_map(h->mTsum, h->Tbef, h->Taft, h->Tincre, h->Ttobe);  // LM:465
  float* Tt = h->Ttobe;
  MapT mt;
  for (int i = 0; i < 6; i++) mt.t[i] = Tt[i];
  mt.sc = host_sincos3(Tt);
  const int ncl = h->n_corner_last, nsl = h->n_surf_last;
  LG_CHECK(h->stack2_c.ensure((size_t)(ncl + 16) * 16, h->st));
  LG_CHECK(h->stack2_s.ensure((size_t)(nsl + 16) * 16, h->st));
  LG_CHECK(h->stack_c.ensure((size_t)(ncl + 16) * 16, h->st));
  LG_CHECK(h->stack_s.ensure((size_t)(nsl + 16) * 16, h->st));
  int rc = lg_map_stack_launch(mt, h->corner_last.as<float4>(), h->stack2_c.as<float4>(), ncl, h->surf_last.as<float4>(),
                               h->stack2_s.as<float4>(), nsl, h->st, &h->launches);
  if (rc) return rc;
  ht.lap(&h->host_s[11]);
  // LM:736-747 down-sample the stacks.  Done ahead of the cube bookkeeping (it needs neither the cubes nor their scratch): the
  // per-cube voxel grids of the previous run may still be running on the second stream and are only waited for below.
  {
    std::vector<VoxSegD> segs(2);
    segs[0] = VoxSegD{h->stack2_c.as<float4>(), nullptr, h->stack_c.as<float4>(), nullptr, ncl, 0.2f};
    segs[1] = VoxSegD{h->stack2_s.as<float4>(), nullptr, h->stack_s.as<float4>(), nullptr, nsl, 0.4f};
    std::vector<int> cnt;
    rc = voxel_segments(h, segs, cnt);
    if (rc) return rc;
    h->n_stack_c = cnt[0];
    h->n_stack_s = cnt[1];
  }
  ht.lap(&h->host_s[HT_MAP_PREP]);
  rc = finish_cube_ds(h);
  if (rc) return rc;

  ht.lap(&h->host_s[9]);
  float yax[3] = {0.f, 10.f, 0.f}, pOnY[3];
  lgh::associate_to_map(Tt, yax, pOnY);  // LM:483-487
  int cI = int((Tt[3] + 25.0) / 50.0) + h->cenW;
  int cJ = int((Tt[4] + 25.0) / 50.0) + h->cenH;
  int cK = int((Tt[5] + 25.0) / 50.0) + h->cenD;
  if (Tt[3] + 25.0 < 0) cI--;
  if (Tt[4] + 25.0 < 0) cJ--;
  if (Tt[5] + 25.0 < 0) cK--;
  while (cI < 3) { cube_shift(h, 0, +1); cI++; h->cenW++; }
  while (cI >= CW - 3) { cube_shift(h, 0, -1); cI--; h->cenW--; }
  while (cJ < 3) { cube_shift(h, 1, +1); cJ++; h->cenH++; }
  while (cJ >= CH - 3) { cube_shift(h, 1, -1); cJ--; h->cenH--; }
  while (cK < 3) { cube_shift(h, 2, +1); cK++; h->cenD++; }
  while (cK >= CD - 3) { cube_shift(h, 2, -1); cK--; h->cenD--; }

  std::vector<int> validInd, surroundInd;  // LM:659-715
  for (int i = cI - 2; i <= cI + 2; i++)
    for (int j = cJ - 2; j <= cJ + 2; j++)
      for (int k = cK - 2; k <= cK + 2; k++) {
        if (i >= 0 && i < CW && j >= 0 && j < CH && k >= 0 && k < CD) {
          float centerX = (float)(50.0 * (i - h->cenW));
          float centerY = (float)(50.0 * (j - h->cenH));
          float centerZ = (float)(50.0 * (k - h->cenD));
          bool inFOV = false;
          for (int ii = -1; ii <= 1; ii += 2)
            for (int jj = -1; jj <= 1; jj += 2)
              for (int kk = -1; kk <= 1; kk += 2) {
                float cornerX = (float)(centerX + 25.0 * ii);
                float cornerY = (float)(centerY + 25.0 * jj);
                float cornerZ = (float)(centerZ + 25.0 * kk);
                float s1 = (Tt[3] - cornerX) * (Tt[3] - cornerX) + (Tt[4] - cornerY) * (Tt[4] - cornerY) + (Tt[5] - cornerZ) * (Tt[5] - cornerZ);
                float s2 = (pOnY[0] - cornerX) * (pOnY[0] - cornerX) + (pOnY[1] - cornerY) * (pOnY[1] - cornerY) +
                           (pOnY[2] - cornerZ) * (pOnY[2] - cornerZ);
                float check1 = (float)(100.0 + s1 - s2 - 10.0 * sqrt(3.0) * sqrtf(s1));
                float check2 = (float)(100.0 + s1 - s2 + 10.0 * sqrt(3.0) * sqrtf(s1));
                if (check1 < 0 && check2 > 0) inFOV = true;
              }
          int ind = i + CW * j + CW * CH * k;
          if (inFOV) validInd.push_back(ind);
          surroundInd.push_back(ind);
        }
      }

  ht.lap(&h->host_s[10]);
  // LM:717-724 gather the local map (reference order: valid cubes in loop order, points in cube order)
  rc = gather_cubes(h, validInd, true, false, h->map_c, &h->n_map_c);
  if (rc) return rc;
  rc = gather_cubes(h, validInd, false, true, h->map_s, &h->n_map_s);
  if (rc) return rc;
  out->n_corner_stack = h->n_stack_c;
  out->n_surf_stack = h->n_stack_s;
  out->n_corner_map = h->n_map_c;
  out->n_surf_map = h->n_map_s;

  if (h->n_map_c > 10 && h->n_map_s > 100) {  // LM:749
    out->optimised = 1;
    ht.lap(&h->host_s[HT_MAP_GRID]);
    {
      int loc[3], hic[3], los[3], his[3];
      cube_box(h, h->cubeC, validInd, loc, hic);
      cube_box(h, h->cubeS, validInd, los, his);
      rc = lg_csr_build2(h->csr, h->map_c.as<float4>(), h->n_map_c, loc, hic, h->map_s.as<float4>(), h->n_map_s, los, his, h->st, &h->launches);
      if (rc) return rc;
    }
    h->grids_valid = true;
    ht.lap(&h->host_s[HT_MAP_ITERS]);
    rc = map_optimize(h, Tt, 10, &out->iterations, false);
    if (rc) return rc;
    for (int i = 0; i < 6; i++) {  // transformUpdate LM:238-241
      h->Tbef[i] = h->mTsum[i];
      h->Taft[i] = Tt[i];
    }
  }
  for (int i = 0; i < 6; i++) mt.t[i] = Tt[i];
  mt.sc = host_sincos3(Tt);

  ht.lap(&h->host_s[HT_MAP_INSERT]);
  // LM:1023-1059 insert the stacks into their cubes
  const int nins = h->n_stack_c + h->n_stack_s;
  std::map<int, std::pair<int, int>> runC, runS;  // cube -> (start in ins_sorted, count)
  if (nins > 0) {
    rc = lg_radix_ensure(h->vb.rs, nins, h->st);
    if (rc) return rc;
    LG_CHECK(h->ins_sel.ensure((size_t)(nins + 16) * 16, h->st));
    LG_CHECK(h->ins_sorted.ensure((size_t)(nins + 16) * 16, h->st));
    const int cap_runs = 2048;
    LG_CHECK(h->d_runs.ensure((size_t)cap_runs * 8 + 16, h->st));
    CubeGeom cg{CW, CH, CD, h->cenW, h->cenH, h->cenD};
    rc = lg_map_insert_launch(mt, cg, h->stack_c.as<float4>(), h->n_stack_c, h->stack_s.as<float4>(), h->n_stack_s, h->ins_sel.as<float4>(),
                              h->vb.rs.keysA.as<unsigned long long>(), h->vb.rs.valsA.as<unsigned int>(), h->st, &h->launches);
    if (rc) return rc;
    int in_b = 0;
    rc = lg_radix_sort(h->vb.rs, nins, 16, h->st, &h->launches, &in_b);
    if (rc) return rc;
    int* d_nruns = h->d_runs.as<int>();
    int2* d_runs = (int2*)(d_nruns + 2);
    rc = lg_map_runs_launch(in_b ? h->vb.rs.keysB.as<unsigned long long>() : h->vb.rs.keysA.as<unsigned long long>(),
                            in_b ? h->vb.rs.valsB.as<unsigned int>() : h->vb.rs.valsA.as<unsigned int>(), h->ins_sel.as<float4>(), nins,
                            h->ins_sorted.as<float4>(), d_nruns, d_runs, cap_runs, h->st, &h->launches);
    if (rc) return rc;
    LG_D2H(h, h->h_ints, d_nruns, 8 + (size_t)cap_runs * 8);
    LG_SYNC(h);
    int nruns = h->h_ints[0];
    if (nruns > cap_runs) return LOAM_ENOSPC;
    std::vector<std::pair<int, int>> runs(nruns);  // (start, key)
    for (int r = 0; r < nruns; r++) runs[r] = {h->h_ints[2 + 2 * r + 1], h->h_ints[2 + 2 * r]};
    std::sort(runs.begin(), runs.end());
    for (int r = 0; r < nruns; r++) {
      int start = runs[r].first, key = runs[r].second;
      int cnt = (r + 1 < nruns ? runs[r + 1].first : nins) - start;
      int cube = key & 0x3fff;
      if (cube == 0x3fff) continue;  // outside the 21 x 11 x 21 grid: dropped (LM:1034-1036)
      if (key & (1 << 14)) runS[cube] = {start, cnt}; else runC[cube] = {start, cnt};
    }
  }
  ht.lap(&h->host_s[HT_MAP_CUBEDS]);
  // LM:1061-1079 voxel-grid every valid cube (old points first, then the new ones: push_back order)
  {
    std::vector<CopyEnt> ents;
    std::vector<int> seg_off(1, 0);
    std::vector<float> leaf;
    const float4* ins = h->ins_sorted.as<float4>();
    int max_n = 0;
    size_t extra = 0;  // raw appends to cubes outside the valid set
    for (auto& kv : runC)
      if (std::find(validInd.begin(), validInd.end(), kv.first) == validInd.end()) extra += kv.second.second;
    for (auto& kv : runS)
      if (std::find(validInd.begin(), validInd.end(), kv.first) == validInd.end()) extra += kv.second.second;
    size_t M = 0;
    for (int ind : validInd) {
      for (auto& c : h->cubeC[ind]) M += c.n;
      for (auto& c : h->cubeS[ind]) M += c.n;
      auto a = runC.find(ind);
      if (a != runC.end()) M += a->second.second;
      auto b = runS.find(ind);
      if (b != runS.end()) M += b->second.second;
    }
    ht.lap(&h->host_s[12]);
    rc = arena_reserve(h, M + extra);
    if (rc) return rc;
    ht.lap(&h->host_s[13]);
    const float4* ar = h->arena.as<float4>();
    // Two layouts of the same input: segment-major [old chunks, new run] per segment (full sort path) and
    // [all old | all new] (merge path: old clouds are already voxel-gridded, only the new points get sorted).
    std::vector<CopyEnt> ents_m;
    std::vector<int> off_old(1, 0), off_new(1, 0);
    bool merge_ok = true;
    for (int ind : validInd)
      for (int type = 0; type < 2; type++) {
        auto& cube = type == 0 ? h->cubeC[ind] : h->cubeS[ind];
        auto& runs = type == 0 ? runC : runS;
        int off = seg_off.back(), oo = off_old.back(), on = off_new.back();
        int nchunks = 0;
        for (auto& c : cube)
          if (c.n > 0) {
            ents.push_back(CopyEnt{ar + c.off, c.n, off});
            ents_m.push_back(CopyEnt{ar + c.off, c.n, oo});
            off += c.n; oo += c.n;
            max_n = std::max(max_n, c.n);
            nchunks++;
          }
        if (nchunks > 1) merge_ok = false;  // raw appended chunks: the old cloud is not a voxel-grid output
        auto it = runs.find(ind);
        if (it != runs.end()) {
          ents.push_back(CopyEnt{ins + it->second.first, it->second.second, off});
          ents_m.push_back(CopyEnt{ins + it->second.first, it->second.second, -1 - on});  // patched below: M_old + on
          off += it->second.second; on += it->second.second;
          max_n = std::max(max_n, it->second.second);
          runs.erase(it);
        }
        seg_off.push_back(off);
        off_old.push_back(oo);
        off_new.push_back(on);
        leaf.push_back(type == 0 ? 0.2f : 0.4f);
      }
    const int n_old = off_old.back(), n_new = off_new.back();
    for (auto& e : ents_m)
      if (e.dst_off < 0) e.dst_off = n_old + (-1 - e.dst_off);
    // runs left over belong to cubes that are not voxel-gridded this time: append raw (push_back semantics)
    {
      std::vector<CopyEnt> app;
      int amax = 0;
      for (int type = 0; type < 2; type++)
        for (auto& kv : (type == 0 ? runC : runS)) {
          auto& cube = type == 0 ? h->cubeC[kv.first] : h->cubeS[kv.first];
          app.push_back(CopyEnt{ins + kv.second.first, kv.second.second, (int)h->bump});
          cube.push_back(Chunk{(int)h->bump, kv.second.second});
          h->bump += kv.second.second;
          amax = std::max(amax, kv.second.second);
        }
      if (!app.empty()) {
        rc = upload(h, h->d_ents, app.data(), app.size() * sizeof(CopyEnt));
        if (rc) return rc;
        rc = lg_gather(h->d_ents.as<CopyEnt>(), (int)app.size(), amax, h->arena.as<float4>(), h->st, &h->launches);
        if (rc) return rc;
      }
    }
    const int nseg = (int)leaf.size();
    const int Mtot = seg_off.back();
    if (nseg > 0 && Mtot > 0) {
      LG_CHECK(h->ds_in.ensure((size_t)(Mtot + 16) * 16, h->st));
      LG_CHECK(h->d_out_se.ensure((size_t)nseg * 8 + 64, h->st));
      int* d_start = h->d_out_se.as<int>();
      int* d_end = d_start + nseg;
      int* d_flags = d_end + nseg;
      float4* outp = h->arena.as<float4>() + h->bump;
      rc = upload(h, h->d_seg_leaf, leaf.data(), nseg * 4);
      if (rc) return rc;
      if (2 * nseg + 1 > loam_handle::H_INTS) return LOAM_ENOSPC;
      loam_handle::PendingDS& pd = h->pds;
      cudaStream_t ds = h->st2;  // behind everything this run has enqueued so far (insert, sort, runs, raw appends, the leaf table)
      LG_CHECK(cudaEventRecord(h->ev_pre_ds, h->st));
      LG_CHECK(cudaStreamWaitEvent(ds, h->ev_pre_ds, 0));
      pd.nseg = nseg;
      pd.Mtot = Mtot;
      pd.max_n = max_n;
      pd.validInd = validInd;
      pd.seg_off = seg_off;
      pd.merged = merge_ok && h->use_merge_path;
      if (pd.merged) {
        pd.ents = ents;
        pd.leaf = leaf;
        std::vector<int> offs3;  // [old | new | merged] offset tables in one upload
        offs3.insert(offs3.end(), off_old.begin(), off_old.end());
        offs3.insert(offs3.end(), off_new.begin(), off_new.end());
        offs3.insert(offs3.end(), seg_off.begin(), seg_off.end());
        rc = upload_on(h, ds, h->d_ents, ents_m.data(), ents_m.size() * sizeof(CopyEnt));
        if (rc) return rc;
        rc = lg_gather(h->d_ents.as<CopyEnt>(), (int)ents_m.size(), max_n, h->ds_in.as<float4>(), ds, &h->launches);
        if (rc) return rc;
        rc = upload_on(h, ds, h->d_seg_off, offs3.data(), offs3.size() * 4);
        if (rc) return rc;
        const int* d_off = h->d_seg_off.as<int>();
        ht.lap(&h->host_s[14]);
        rc = lg_vox_merge(h->vb, h->ds_in.as<float4>(), d_off, d_off + (nseg + 1), d_off + 2 * (nseg + 1), h->d_seg_leaf.as<float>(), nseg,
                          n_old, n_new, outp, d_start, d_end, d_flags, ds, &h->launches);
        if (rc) return rc;
        ht.lap(&h->host_s[15]);
        h->d2h_bytes += (long long)nseg * 8 + 4;
        LG_CHECK(cudaMemcpyAsync(h->h_ints2, d_start, (size_t)nseg * 8 + 4, cudaMemcpyDeviceToHost, ds));
      } else {
        rc = upload_on(h, ds, h->d_ents, ents.data(), ents.size() * sizeof(CopyEnt));
        if (rc) return rc;
        rc = lg_gather(h->d_ents.as<CopyEnt>(), (int)ents.size(), max_n, h->ds_in.as<float4>(), ds, &h->launches);
        if (rc) return rc;
        rc = upload_on(h, ds, h->d_seg_off, seg_off.data(), (nseg + 1) * 4);
        if (rc) return rc;
        ht.lap(&h->host_s[14]);
        rc = lg_vox_big(h->vb, h->ds_in.as<float4>(), h->d_seg_off.as<int>(), h->d_seg_leaf.as<float>(), nseg, Mtot, outp, d_start, d_end, ds,
                        &h->launches);
        if (rc) return rc;
        ht.lap(&h->host_s[15]);
        h->d2h_bytes += (long long)nseg * 8;
        LG_CHECK(cudaMemcpyAsync(h->h_ints2, d_start, (size_t)nseg * 8, cudaMemcpyDeviceToHost, ds));
      }
      LG_CHECK(cudaEventRecord(pd.ev, ds));
      pd.active = true;
      // this run's surround cloud (LM:1081-1101) gathers the cubes just written: it needs the descriptors now
      if (h->prm.want_surround && h->mapFrameCount + 1 >= 5) {
        rc = finish_cube_ds(h);
        if (rc) return rc;
      }
    }
  }
  ht.lap(&h->host_s[HT_MAP_REST]);
  // LM:1081-1101
  h->mapFrameCount++;
  if (h->mapFrameCount >= 5) {
    h->mapFrameCount = 0;
    out->surround_published = 1;
    if (h->prm.want_surround && h->aux && h->aux_reserved) {
      // pipelined: enqueue the gather of the 125 cubes on the output stream behind this run's map kernels and hand the rest
      // (voxel grid 0.2 m + count) to the output thread; out->n_surround is filled in there
      loam_handle* a = h->aux;
      std::vector<CopyEnt> ents;
      const float4* ar = h->arena.as<float4>();
      int off = 0, max_n = 0;
      for (int ind : surroundInd) {
        for (auto& c : h->cubeC[ind])
          if (c.n > 0) { ents.push_back(CopyEnt{ar + c.off, c.n, off}); off += c.n; max_n = std::max(max_n, c.n); }
        for (auto& c : h->cubeS[ind])
          if (c.n > 0) { ents.push_back(CopyEnt{ar + c.off, c.n, off}); off += c.n; max_n = std::max(max_n, c.n); }
      }
      LG_CHECK(cudaEventRecord(h->ev_map_done, h->st));
      LG_CHECK(cudaStreamWaitEvent(a->st, h->ev_map_done, 0));
      LG_CHECK(a->vg_in.ensure((size_t)(off + 16) * 16, a->st));
      if (!ents.empty()) {
        rc = upload(a, a->d_ents, ents.data(), ents.size() * sizeof(CopyEnt));  // pageable source: staged before the call returns
        if (rc) return rc;
        rc = lg_gather(a->d_ents.as<CopyEnt>(), (int)ents.size(), max_n, a->vg_in.as<float4>(), a->st, &a->launches);
        if (rc) return rc;
      }
      LG_CHECK(cudaEventRecord(h->ev_aux_read, a->st));
      h->aux_read_pending = true;
      h->aux_ns = off;
      out->n_surround = -1;
    } else if (h->prm.want_surround) {
      DevBuf& tmp = h->vg_in;
      int ns = 0;
      rc = gather_cubes(h, surroundInd, true, true, tmp, &ns);
      if (rc) return rc;
      LG_CHECK(h->surround.ensure((size_t)(ns + 16) * 16, h->st));
      std::vector<VoxSegD> segs(1);
      segs[0] = VoxSegD{tmp.as<float4>(), nullptr, h->surround.as<float4>(), nullptr, ns, 0.2f};
      std::vector<int> cnt;
      rc = voxel_segments(h, segs, cnt);
      if (rc) return rc;
      h->n_surround = cnt[0];
      out->n_surround = cnt[0];
    }
  }
  // LM:1103-1106
  if (h->prm.want_registered) {
    LG_CHECK(h->registered.ensure((size_t)(h->n_fullres3 + 16) * 16, h->st));
    rc = lg_map_register_launch(mt, h->fullres3.as<float4>(), h->registered.as<float4>(), h->n_fullres3, h->st, &h->launches);
    if (rc) return rc;
    h->n_registered = h->n_fullres3;
    out->n_registered = h->n_registered;
  }
  for (int i = 0; i < 6; i++) {
    out->transform_aft_mapped[i] = h->Taft[i];
    out->transform_bef_mapped[i] = h->Tbef[i];
    out->transform_tobe_mapped[i] = Tt[i];
  }
  return LOAM_OK;
}

// ============================================================================================ transformMaintenance
int loam_integrate_odometry(loam_handle* h, const float* Tsum, double stamp, float* out6, double* track4) {
  if (!h || !Tsum || !out6 || !track4) return LOAM_EINVAL;
  if (fabs((double)Tsum[3]) < 0.000001 && fabs((double)Tsum[4]) < 0.000001 && fabs((double)Tsum[5]) < 0.000001) {  // TM:264-275
    h->tmPre[3] = 0;
    for (int i = 0; i < 6; i++) h->tmSum[i] = h->tmIncre[i] = h->tmMapped[i] = h->tmBef[i] = h->tmAft[i] = 0;
  }
  for (int i = 0; i < 6; i++) h->tmSum[i] = Tsum[i];
  lgh::transform_associate_to_map(h->tmSum, h->tmBef, h->tmAft, h->tmIncre, h->tmMapped);  // TM:175-260
  for (int i = 0; i < 6; i++) out6[i] = h->tmMapped[i];
  double px = h->tmMapped[3], py = h->tmMapped[4], pz = h->tmMapped[5];  // TM:116-157
  double* pre = h->tmPre;
  double* tmp = h->tmTmp;
  if (pre[3] == 0) {
    pre[0] = pz; pre[1] = px; pre[2] = py; pre[3] = stamp;
    for (int i = 0; i < 4; i++) tmp[i] = pre[i];
  } else {
    double dX = pz - pre[0], dY = px - pre[1], dZ = py - pre[2];
    double dX1 = dX * sqrt(pow(dX, 2) + pow(dY, 2) + pow(dZ, 2)) / sqrt(pow(dX, 2) + pow(dY, 2));
    double dY1 = dY * sqrt(pow(dX, 2) + pow(dY, 2) + pow(dZ, 2)) / sqrt(pow(dX, 2) + pow(dY, 2));
    tmp[0] += dX1; tmp[1] += dY1; tmp[2] = py; tmp[3] = stamp;
    pre[0] = pz; pre[1] = px; pre[2] = py; pre[3] = stamp;
  }
  track4[0] = tmp[0]; track4[1] = tmp[1]; track4[2] = 10; track4[3] = tmp[3];
  return LOAM_OK;
}
int loam_integrate_mapping(loam_handle* h, const float* aft, const float* bef) {
  if (!h || !aft || !bef) return LOAM_EINVAL;
  for (int i = 0; i < 6; i++) { h->tmAft[i] = aft[i]; h->tmBef[i] = bef[i]; }
  return LOAM_OK;
}

// ============================================================================================ whole sweep
static int process_common(loam_handle* h, loam_sweep_result* out) {
  int rc = loam_odometry_process(h, &out->odom);
  if (rc) return rc;
  out->mapping_ran = 0;
  if (out->odom.odom_published) {
    float hop[6];
    if (h->prm.pose_message_hop) lgh::pose_message_hop(out->odom.transform_sum, hop);
    rc = loam_mapping_odometry(h, h->prm.pose_message_hop ? hop : out->odom.transform_sum);
    if (rc) return rc;
  }
  if (out->odom.odom_published && out->odom.fullres_published) {
    rc = loam_mapping_process(h, &out->map);
    if (rc) return rc;
    out->mapping_ran = 1;
  }
  return LOAM_OK;
}
int loam_process_sweep(loam_handle* h, const float* xyz_host, int n, int stride_bytes, double stamp, loam_sweep_result* out) {
  if (!h || !out) return LOAM_EINVAL;
  memset(out, 0, sizeof(*out));
  int rc = loam_extract(h, xyz_host, n, stride_bytes, stamp, nullptr, &out->counts);
  if (rc) return rc;
  return process_common(h, out);
}
int loam_process_sweep_device(loam_handle* h, const float* xyz_dev, int n, int stride_bytes, double stamp, loam_sweep_result* out) {
  if (!h || !out) return LOAM_EINVAL;
  memset(out, 0, sizeof(*out));
  int rc = loam_extract_device(h, xyz_dev, n, stride_bytes, stamp, nullptr, &out->counts);
  if (rc) return rc;
  return process_common(h, out);
}

// ============================================================================================ data access
static int select_cloud(loam_handle* h, int which, const void** src, int* cnt) {
  switch (which) {
    case LOAM_CLOUD_FULL: *src = h->cur_full; *cnt = h->counts.n_full; break;
    case LOAM_CLOUD_SHARP: *src = h->cur_sharp; *cnt = h->counts.n_sharp; break;
    case LOAM_CLOUD_LESS_SHARP: *src = h->cur_less_sharp; *cnt = h->counts.n_less_sharp; break;
    case LOAM_CLOUD_FLAT: *src = h->cur_flat; *cnt = h->counts.n_flat; break;
    case LOAM_CLOUD_LESS_FLAT: *src = h->cur_less_flat; *cnt = h->counts.n_less_flat; break;
    case LOAM_CLOUD_CORNER_LAST: *src = h->corner_last.p; *cnt = h->n_corner_last; break;
    case LOAM_CLOUD_SURF_LAST: *src = h->surf_last.p; *cnt = h->n_surf_last; break;
    case LOAM_CLOUD_FULL_RES3: *src = h->fullres3.p; *cnt = h->n_fullres3; break;
    case LOAM_CLOUD_CORNER_STACK: *src = h->stack_c.p; *cnt = h->n_stack_c; break;
    case LOAM_CLOUD_SURF_STACK: *src = h->stack_s.p; *cnt = h->n_stack_s; break;
    case LOAM_CLOUD_CORNER_MAP: *src = h->map_c.p; *cnt = h->n_map_c; break;
    case LOAM_CLOUD_SURF_MAP: *src = h->map_s.p; *cnt = h->n_map_s; break;
    case LOAM_CLOUD_SURROUND: *src = h->surround.p; *cnt = h->n_surround; break;
    case LOAM_CLOUD_REGISTERED: *src = h->registered.p; *cnt = h->n_registered; break;
    default: return LOAM_EINVAL;
  }
  return LOAM_OK;
}

int loam_get_cloud(loam_handle* h, int which, float* host_buf, int cap, int* n) {
  if (!h || !n) return LOAM_EINVAL;
  LG_CHECK(cudaSetDevice(h->device));
  g_lg_prof = h->prof.on ? &h->prof : nullptr;
  const void* src = nullptr;
  int cnt = 0;
  int rc = select_cloud(h, which, &src, &cnt);
  if (rc) return rc;
  *n = cnt;
  if (!host_buf) return LOAM_OK;
  if (cap < cnt) return LOAM_ENOSPC;
  if (cnt > 0) {
    LG_D2H(h, host_buf, src, (size_t)cnt * 16);
    LG_SYNC(h);
  }
  return LOAM_OK;
}

int loam_get_cloud_wire(loam_handle* h, int which, void* host_buf, int cap_points, int* n_points) {
  if (!h || !n_points) return LOAM_EINVAL;
  LG_CHECK(cudaSetDevice(h->device));
  g_lg_prof = h->prof.on ? &h->prof : nullptr;
  const void* src = nullptr;
  int cnt = 0;
  int rc = select_cloud(h, which, &src, &cnt);
  if (rc) return rc;
  *n_points = cnt;
  if (!host_buf) return LOAM_OK;
  if (cap_points < cnt) return LOAM_ENOSPC;
  if (cnt > 0) {
    LG_CHECK(h->wire.ensure((size_t)cnt * 32 + 16, h->st));
    pack_pcl32_kernel<<<lg_div_up(cnt, 256), 256, 0, h->st>>>((const float4*)src, cnt, h->wire.as<float4>());
    h->launches++;
    LG_CHECK(cudaGetLastError());
    LG_D2H(h, host_buf, h->wire.p, (size_t)cnt * 32);
    LG_SYNC(h);
  }
  return LOAM_OK;
}

int loam_get_diag(loam_handle* h, int which, void* host_buf, int cap_bytes, int* n_items) {
  if (!h || !n_items) return LOAM_EINVAL;
  if (!h->sr.meta.p) return LOAM_ESTATE;
  LG_CHECK(cudaSetDevice(h->device));
  g_lg_prof = h->prof.on ? &h->prof : nullptr;
  const void* src = nullptr;
  int cnt = 0, esz = 1;
  switch (which) {
    case LOAM_DIAG_CURVATURE: src = h->sr.curv.p; cnt = h->counts.n_full; esz = 4; break;
    case LOAM_DIAG_PICKED_MASK: src = h->sr.mask_diag.p; cnt = h->counts.n_full; esz = 1; break;
    case LOAM_DIAG_LABEL: src = h->sr.label.p; cnt = h->counts.n_full; esz = 1; break;
    case LOAM_DIAG_SCAN_START: src = h->sr.meta.as<int>() + SRM_SCAN_START; cnt = h->prm.n_scans; esz = 4; break;
    case LOAM_DIAG_SCAN_END: src = h->sr.meta.as<int>() + SRM_SCAN_END; cnt = h->prm.n_scans; esz = 4; break;
    default: return LOAM_EINVAL;
  }
  *n_items = cnt;
  if (!host_buf) return LOAM_OK;
  if (cap_bytes < cnt * esz) return LOAM_ENOSPC;
  if (cnt > 0) {
    LG_D2H(h, host_buf, src, (size_t)cnt * esz);
    LG_SYNC(h);
  }
  return LOAM_OK;
}

// ============================================================================================ stage-level
int loam_voxel_grid(loam_handle* h, const float* in4_host, int m, float leaf, float* out4_host, int cap, int* v) {
  if (!h || !v || m < 0 || !(leaf > 0.f)) return LOAM_EINVAL;
  LG_CHECK(cudaSetDevice(h->device));
  g_lg_prof = h->prof.on ? &h->prof : nullptr;
  *v = 0;
  if (m == 0) return LOAM_OK;
  int rc = upload(h, h->vg_in, in4_host, (size_t)m * 16);
  if (rc) return rc;
  DevBuf outb;
  LG_CHECK(outb.ensure((size_t)(m + 16) * 16, h->st));
  std::vector<VoxSegD> segs(1);
  segs[0] = VoxSegD{h->vg_in.as<float4>(), nullptr, outb.as<float4>(), nullptr, m, leaf};
  std::vector<int> cnt;
  rc = voxel_segments(h, segs, cnt);
  if (rc == LOAM_OK) {
    *v = cnt[0];
    if (out4_host) {
      if (cap < cnt[0]) rc = LOAM_ENOSPC;
      else if (cnt[0] > 0) {
        cudaError_t e = cudaMemcpyAsync(out4_host, outb.p, (size_t)cnt[0] * 16, cudaMemcpyDeviceToHost, h->st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->st);
        if (e != cudaSuccess) { lg_set_error(cudaGetErrorString(e), __FILE__, __LINE__); rc = LOAM_ECUDA; }
      }
    }
  }
  cudaStreamSynchronize(h->st);
  outb.release();
  return rc;
}

int loam_odom_set_inputs(loam_handle* h, const float* sharp, int n_sharp, const float* flat, int n_flat, const float* corner_last,
                         int n_corner_last, const float* surf_last, int n_surf_last) {
  if (!h || n_sharp < 0 || n_flat < 0 || n_corner_last < 0 || n_surf_last < 0) return LOAM_EINVAL;
  LG_CHECK(cudaSetDevice(h->device));
  g_lg_prof = h->prof.on ? &h->prof : nullptr;
  int rc = upload(h, h->t_sharp, sharp, (size_t)n_sharp * 16);
  if (!rc) rc = upload(h, h->t_flat, flat, (size_t)n_flat * 16);
  if (!rc) rc = upload(h, h->corner_last, corner_last, (size_t)n_corner_last * 16);
  if (!rc) rc = upload(h, h->surf_last, surf_last, (size_t)n_surf_last * 16);
  if (rc) return rc;
  LG_SYNC(h);
  h->cur_sharp = h->t_sharp.as<float4>();
  h->cur_flat = h->t_flat.as<float4>();
  h->counts.n_sharp = n_sharp;
  h->counts.n_flat = n_flat;
  h->n_corner_last = n_corner_last;
  h->od.bounds_valid = false;
  h->n_surf_last = n_surf_last;
  return LOAM_OK;
}

int loam_odom_iter(loam_handle* h, int iter, const float* T, float* AtA, float* AtB, int* n_sel) {
  if (!h || !T || !AtA || !AtB || !n_sel || iter < 0) return LOAM_EINVAL;
  if (!h->cur_sharp && h->counts.n_sharp > 0) return LOAM_ESTATE;
  LG_CHECK(cudaSetDevice(h->device));
  g_lg_prof = h->prof.on ? &h->prof : nullptr;
  int rc = odom_iter(h, iter, T, AtA, AtB, n_sel);
  if (rc) return rc;
  if (*n_sel < 10) {
    memset(AtA, 0, 36 * sizeof(float));
    memset(AtB, 0, 6 * sizeof(float));
  }
  return LOAM_OK;
}

int loam_odom_get_corr(loam_handle* h, int* c1, int* c2, int cap_c, int* s1, int* s2, int* s3, int cap_s) {
  if (!h) return LOAM_EINVAL;
  if (cap_c < h->counts.n_sharp || cap_s < h->counts.n_flat) return LOAM_ENOSPC;
  LG_CHECK(cudaSetDevice(h->device));
  g_lg_prof = h->prof.on ? &h->prof : nullptr;
  size_t bc = (size_t)h->counts.n_sharp * 4, bs = (size_t)h->counts.n_flat * 4;
  if (bc) {
    LG_D2H(h, c1, h->od.c1.p, bc);
    LG_D2H(h, c2, h->od.c2.p, bc);
  }
  if (bs) {
    LG_D2H(h, s1, h->od.s1.p, bs);
    LG_D2H(h, s2, h->od.s2.p, bs);
    LG_D2H(h, s3, h->od.s3.p, bs);
  }
  LG_SYNC(h);
  return LOAM_OK;
}

int loam_transform_to_end(loam_handle* h, const float* in4_host, int n, const float* T, const float* imu_trans, float* out4_host) {
  if (!h || n < 0 || !T) return LOAM_EINVAL;
  LG_CHECK(cudaSetDevice(h->device));
  g_lg_prof = h->prof.on ? &h->prof : nullptr;
  if (n == 0) return LOAM_OK;
  float imu[12] = {0};
  if (imu_trans) memcpy(imu, imu_trans, sizeof(imu));
  int rc = upload(h, h->vg_in, in4_host, (size_t)n * 16);
  if (rc) return rc;
  LG_CHECK(h->vg_out.ensure((size_t)(n + 16) * 16, h->st));
  OdomT ot;
  for (int i = 0; i < 6; i++) ot.t[i] = T[i];
  rc = lg_odom_to_end_launch(ot, host_sincos3(T), imu_sc(imu), h->vg_in.as<float4>(), h->vg_out.as<float4>(), n, nullptr, nullptr, 0, nullptr,
                             nullptr, 0, h->st, &h->launches);
  if (rc) return rc;
  LG_D2H(h, out4_host, h->vg_out.p, (size_t)n * 16);
  LG_SYNC(h);
  return LOAM_OK;
}

int loam_map_set_inputs(loam_handle* h, const float* corner_stack, int n_cs, const float* surf_stack, int n_ss, const float* corner_map, int n_cm,
                        const float* surf_map, int n_sm) {
  if (!h || n_cs < 0 || n_ss < 0 || n_cm < 0 || n_sm < 0) return LOAM_EINVAL;
  LG_CHECK(cudaSetDevice(h->device));
  g_lg_prof = h->prof.on ? &h->prof : nullptr;
  int rc = upload(h, h->stack_c, corner_stack, (size_t)n_cs * 16);
  if (!rc) rc = upload(h, h->stack_s, surf_stack, (size_t)n_ss * 16);
  if (!rc) rc = upload(h, h->map_c, corner_map, (size_t)n_cm * 16);
  if (!rc) rc = upload(h, h->map_s, surf_map, (size_t)n_sm * 16);
  if (rc) return rc;
  h->n_stack_c = n_cs; h->n_stack_s = n_ss; h->n_map_c = n_cm; h->n_map_s = n_sm;
  h->lm_gn = LgGNState();  // a new explicit problem starts with the reference's initial matP / isDegenerate (LM:399-400)
  // explicit clouds: their boxes come from the data (the mapping node knows them from its cubes)
  LG_CHECK(h->csr.bb.ensure(64, h->st));
  rc = lg_csr_bbox_launch(h->map_c.as<float4>(), n_cm, h->map_s.as<float4>(), n_sm, h->csr.bb.as<int>(), h->st, &h->launches);
  if (rc) return rc;
  LG_D2H(h, h->h_ints, h->csr.bb.p, 12 * 4);
  LG_SYNC(h);
  int lo[2][3], hi[2][3];
  for (int g = 0; g < 2; g++)
    for (int a = 0; a < 3; a++) {
      lo[g][a] = h->h_ints[g * 6 + a];
      hi[g][a] = h->h_ints[g * 6 + 3 + a];
      if (lo[g][a] > hi[g][a]) lo[g][a] = hi[g][a] = 0;  // empty cloud
    }
  rc = lg_csr_build2(h->csr, h->map_c.as<float4>(), n_cm, lo[0], hi[0], h->map_s.as<float4>(), n_sm, lo[1], hi[1], h->st, &h->launches);
  if (rc) return rc;
  LG_SYNC(h);
  h->grids_valid = true;
  return LOAM_OK;
}

int loam_map_iter(loam_handle* h, int iter, const float* T, float* AtA, float* AtB, int* n_sel) {
  if (!h || !T || !AtA || !AtB || !n_sel || iter < 0) return LOAM_EINVAL;
  if (!h->grids_valid) return LOAM_ESTATE;
  LG_CHECK(cudaSetDevice(h->device));
  g_lg_prof = h->prof.on ? &h->prof : nullptr;
  int rc = map_iter(h, iter, T, AtA, AtB, n_sel);
  if (rc) return rc;
  if (*n_sel < 50) {
    memset(AtA, 0, 36 * sizeof(float));
    memset(AtB, 0, 6 * sizeof(float));
  }
  return LOAM_OK;
}

int loam_map_get_corr(loam_handle* h, int* corner5, int cap_c, int* surf5, int cap_s) {
  if (!h) return LOAM_EINVAL;
  if (cap_c < h->n_stack_c || cap_s < h->n_stack_s) return LOAM_ENOSPC;
  if (!h->gn.nbr.p || !h->nbr_valid) return LOAM_ESTATE;  // only after a stage-level iteration (loam_map_iter*)
  LG_CHECK(cudaSetDevice(h->device));
  g_lg_prof = h->prof.on ? &h->prof : nullptr;
  if (h->n_stack_c) LG_D2H(h, corner5, h->gn.nbr.p, (size_t)h->n_stack_c * 20);
  if (h->n_stack_s)
    LG_D2H(h, surf5, h->gn.nbr.as<int>() + (size_t)h->n_stack_c * 5, (size_t)h->n_stack_s * 20);
  LG_SYNC(h);
  return LOAM_OK;
}

int loam_gn_solve(const float* AtA, const float* AtB, int iter, float eig_threshold, float* state37, float* X) {
  if (!AtA || !AtB || !state37 || !X) return LOAM_EINVAL;
  LgGNState st;
  memcpy(st.matP, state37, sizeof(st.matP));
  st.degenerate = state37[36] != 0.f;
  lg_gn_solve_step(AtA, AtB, iter, eig_threshold, st, X);
  memcpy(state37, st.matP, sizeof(st.matP));
  state37[36] = st.degenerate ? 1.f : 0.f;
  return LOAM_OK;
}

int loam_map_iter_partial(loam_handle* h, int iter, const float* T, double* partial_dev28) {
  if (!h || !T || !partial_dev28 || iter < 0) return LOAM_EINVAL;
  if (!h->grids_valid) return LOAM_ESTATE;
  LG_CHECK(cudaSetDevice(h->device));
  g_lg_prof = h->prof.on ? &h->prof : nullptr;
  int rc = map_gn_enqueue(h, T, iter, iter + 1, 0, partial_dev28, false);
  if (rc) return rc;
  LG_SYNC(h);
  return LOAM_OK;
}

// ---- sharded map, fused all-reduce over NVLink peer memory (SURVEY 8e "optimised") -------------------------------------
int loam_shard_export(loam_handle* h, unsigned char* handle64) {
  if (!h || !handle64) return LOAM_EINVAL;
  LG_CHECK(cudaSetDevice(h->device));
  if (!h->xchg) {
    LG_CHECK(cudaMalloc((void**)&h->xchg, LG_XCHG_BYTES));
    LG_CHECK(cudaMemset(h->xchg, 0, LG_XCHG_BYTES));
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t hd;
  LG_CHECK(cudaIpcGetMemHandle(&hd, h->xchg));
  memcpy(handle64, &hd, 64);
  {
    std::lock_guard<std::mutex> lock(g_xchg_mutex);
    g_xchg_local[std::string((const char*)handle64, 64)] = h->xchg;
  }
  return LOAM_OK;
}

int loam_shard_connect(loam_handle* h, const unsigned char* handles, int world, int rank) {
  if (!h || !handles || world < 1 || world > LG_MAX_PEERS || rank < 0 || rank >= world) return LOAM_EINVAL;
  if (!h->xchg || h->px_connected) return LOAM_ESTATE;
  LG_CHECK(cudaSetDevice(h->device));
  memset(&h->px, 0, sizeof(h->px));
  h->px.world = world;
  h->px.rank = rank;
  h->h_mail[28] = 0.0;
  h->px.timeout = (int*)(h->d_mail + 28);  // travels with the sums through the mapped mailbox
  for (int r = 0; r < world; r++) {
    if (r == rank) {
      h->px.buf[r] = h->xchg;
      continue;
    }
    {
      std::lock_guard<std::mutex> lock(g_xchg_mutex);
      auto it = g_xchg_local.find(std::string((const char*)handles + (size_t)r * 64, 64));
      if (it != g_xchg_local.end()) {
        h->px.buf[r] = it->second;
        h->px_local[r] = true;
        continue;
      }
    }
    cudaIpcMemHandle_t hd;
    memcpy(&hd, handles + (size_t)r * 64, 64);
    void* p = nullptr;
    LG_CHECK(cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
    h->px.buf[r] = (double*)p;
  }
  h->px_connected = true;
  return LOAM_OK;
}

int loam_map_iter_allreduce(loam_handle* h, int iter, const float* T, float* AtA, float* AtB, int* n_sel) {
  if (!h || !T || !AtA || !AtB || !n_sel || iter < 0) return LOAM_EINVAL;
  if (!h->grids_valid || !h->px_connected) return LOAM_ESTATE;
  LG_CHECK(cudaSetDevice(h->device));
  g_lg_prof = h->prof.on ? &h->prof : nullptr;
  int rc = map_iter(h, iter, T, AtA, AtB, n_sel, true);
  if (rc) return rc;
  if (*n_sel < 50) {  // LM:929-932 on the global count
    memset(AtA, 0, 36 * sizeof(float));
    memset(AtB, 0, 6 * sizeof(float));
  }
  if (*(volatile int*)(h->h_mail + 28)) {  // written by the kernel next to the sums (mapped mailbox)
    lg_set_error("a peer rank never reached the all-reduce", __FILE__, __LINE__);
    return LOAM_ECUDA;
  }
  return LOAM_OK;
}

int loam_shard_inject(loam_handle* h, int from_rank, const double* sums28_host) {
  if (!h || !sums28_host || from_rank < 0 || from_rank >= LG_MAX_PEERS) return LOAM_EINVAL;
  if (!h->px_connected || from_rank == h->px.rank || from_rank >= h->px.world) return LOAM_ESTATE;
  LG_CHECK(cudaSetDevice(h->device));
  LG_CHECK(cudaStreamSynchronize(h->st));
  const unsigned long long xs = h->px.xseq + 1;  // the next iteration's sequence number
  const size_t slot = ((size_t)(xs & 1ull) * LG_MAX_PEERS + from_rank) * 32;
  LG_CHECK(cudaMemcpy(h->xchg + slot, sums28_host, 28 * 8, cudaMemcpyHostToDevice));
  LG_CHECK(cudaMemcpy((unsigned long long*)(h->xchg + LG_XCHG_FLAG_OFFSET) + from_rank, &xs, 8, cudaMemcpyHostToDevice));
  return LOAM_OK;
}

int loam_shard_set_slab(loam_handle* h, float x_lo, float x_hi) {
  if (!h || !(x_lo < x_hi)) return LOAM_EINVAL;
  h->slab_lo = x_lo;
  h->slab_hi = x_hi;
  return LOAM_OK;
}

int loam_map_optimize(loam_handle* h, float* T, int max_iters, int* iterations) {
  if (!h || !T || !iterations || max_iters < 0 || max_iters > 1000) return LOAM_EINVAL;
  if (!h->grids_valid) return LOAM_ESTATE;
  LG_CHECK(cudaSetDevice(h->device));
  g_lg_prof = h->prof.on ? &h->prof : nullptr;
  int rc = map_optimize(h, T, max_iters, iterations, h->px_connected);
  if (rc) return rc;
  if (h->px_connected && *(volatile int*)(h->h_mail + 28)) {
    lg_set_error("a peer rank never reached the all-reduce", __FILE__, __LINE__);
    return LOAM_ECUDA;
  }
  return LOAM_OK;
}

int loam_map_finish_reduced(const double* reduced28_host, float* AtA, float* AtB, int* n_sel) {
  if (!reduced28_host || !AtA || !AtB || !n_sel) return LOAM_EINVAL;
  lg_unpack28(reduced28_host, AtA, AtB, n_sel);
  if (*n_sel < 50) {
    memset(AtA, 0, 36 * sizeof(float));
    memset(AtB, 0, 6 * sizeof(float));
  }
  return LOAM_OK;
}

}  // extern "C"

// ===================================================================================================== pipelined mode
// The reference runs scanRegistration, laserOdometry and laserMapping as three processes connected by ROS queues
// (SURVEY §1): while mapping works on sweep k, odometry registers sweep k+1 and scanRegistration extracts sweep k+2.
// loam_pipeline_* is the same structure on one GPU: three stage threads, each with its own handle (state + CUDA
// stream), device-resident hand-over through small rings of slots ordered by CUDA events.  Results are identical to
// loam_process_sweep (same kernels, same order of operations per stage); only the latency/throughput trade changes.
#include <atomic>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>

namespace {

template <typename T>
struct BQueue {
  std::mutex m;
  std::condition_variable cv;
  std::deque<T> q;
  void push(const T& v) {
    { std::lock_guard<std::mutex> l(m); q.push_back(v); }
    cv.notify_one();
  }
  T pop() {
    std::unique_lock<std::mutex> l(m);
    cv.wait(l, [&] { return !q.empty(); });
    T v = q.front();
    q.pop_front();
    return v;
  }
};
struct Sem {
  std::mutex m;
  std::condition_variable cv;
  int n;
  explicit Sem(int v) : n(v) {}
  void acquire() {
    std::unique_lock<std::mutex> l(m);
    cv.wait(l, [&] { return n > 0; });
    n--;
  }
  void release() {
    { std::lock_guard<std::mutex> l(m); n++; }
    cv.notify_one();
  }
};
enum { JOB_SWEEP = 0, JOB_RESET = 1, JOB_STOP = 2, JOB_IMU = 3 };
struct Job {
  long long k;
  int kind;
  int slot;           // input slot (stage A), feature slot (stage B), map slot or -1 (stage C)
  const float* xyz;   // device pointer of the sweep (stage A)
  int n, stride;
  int odom_published, full;
  float Tsum[6];
  long long epoch;    // number of loam_pipeline_reset calls before this job: an error only poisons its own epoch
  int pre;            // the sweep was extracted by loam_pipeline_submit_batch already: stage A only hands the clouds on
  loam_counts counts;
  double stamp;       // header stamp of the sweep (timeScanCur SR:257) or of the IMU message
  double imu_msg[10]; // JOB_IMU: orientation {x, y, z, w}, angular velocity, linear acceleration
};
constexpr int PNS = 4;  // slots per ring

}  // namespace

struct loam_pipeline {
  loam_handle *hA = nullptr, *hB = nullptr, *hC = nullptr;
  loam_handle* hD = nullptr;  // output handle: /laser_cloud_surround is finished here, off the mapping stage's critical path
  int device = 0;
  cudaStream_t copy_st = nullptr;
  DevBuf in_xyz[PNS];
  cudaEvent_t in_copied[PNS];
  struct Feat {
    DevBuf b[5];
    loam_counts c;
    float imu[12];  // /imu_trans of the sweep (SR:730-745), consumed by the odometry stage (LO:201-225, 566-568, 1053-1064)
    cudaEvent_t ready, consumed;
  } feat[PNS];
  struct MapIn {
    DevBuf corner, surf, full;
    int nc, ns, nf;
    cudaEvent_t ready, consumed;
  } mapin[PNS];
  Sem in_free{PNS}, feat_free{PNS}, map_free{PNS};
  BQueue<Job> qA, qB, qC, qD;
  Sem aux_free{1};  // one surround job in flight on the output handle
  std::mutex rm;
  std::condition_variable rcv;
  std::map<long long, loam_sweep_result> partial, done;
  long long next_submit = 0, next_wait = 0, in_count = 0, feat_count = 0, map_count = 0;
  // first error of the current epoch (an epoch ends at loam_pipeline_reset): sweeps of that epoch are skipped and
  // loam_pipeline_wait reports the code for them; sweeps submitted after the reset run normally again
  std::atomic<int> error{0};
  std::atomic<long long> error_epoch{-1};
  long long epoch = 0;
  std::map<long long, long long> epoch_of;  // sweep -> epoch (for loam_pipeline_wait)
  char err_text[512] = "";
  long long resets_pushed = 0, resets_at_b = 0;  // loam_pipeline_reset calls / those the odometry stage has seen (guarded by rm)
  long long pre_submitted = 0, pre_handled = 0;  // batch-extracted sweeps pushed / handed on by stage A (guarded by rm)
  long long imu_pushed = 0, imu_handled = 0;     // loam_pipeline_imu_push calls / those stage A has applied (guarded by rm)
  double busy[3] = {0, 0, 0};  // seconds each stage thread spent working on sweeps (not waiting for its queue / a free slot)
  std::thread tA, tB, tC, tD;
};

namespace {

void pipe_fail(loam_pipeline* p, int rc, long long epoch) {
  {
    std::lock_guard<std::mutex> l(p->rm);
    if (p->error_epoch.load() != epoch) {  // first error of this epoch (a stale one from an earlier epoch is replaced)
      snprintf(p->err_text, sizeof(p->err_text), "%s", g_cuda_err);  // the stage thread's text, visible to the caller
      p->error.store(rc);
      p->error_epoch.store(epoch);
    }
  }
  p->rcv.notify_all();
}
int pipe_error(loam_pipeline* p, long long epoch) { return p->error_epoch.load() == epoch ? p->error.load() : 0; }

void stage_a(loam_pipeline* p) {
  cudaSetDevice(p->device);
  loam_handle* h = p->hA;
  for (;;) {
    Job j = p->qA.pop();
    if (j.kind == JOB_IMU) {  // imuHandler SR:754-837, in submission order with the sweeps
      loam_imu_push(h, j.stamp, j.imu_msg, j.imu_msg + 4, j.imu_msg + 7);
      {
        std::lock_guard<std::mutex> l(p->rm);
        p->imu_handled++;
      }
      p->rcv.notify_all();
      continue;
    }
    if (j.kind != JOB_SWEEP) {
      p->qB.push(j);
      if (j.kind == JOB_STOP) return;
      continue;
    }
    // the feature slot is taken BEFORE the extraction: the kernels write the five clouds straight into the slot's buffers
    // (swapped into the workspace for the duration of the sweep), so nothing is copied between the two stages
    p->feat_free.acquire();
    const int fs = (int)(p->feat_count++ % PNS);
    loam_pipeline::Feat& f = p->feat[fs];
    const auto t_busy0 = std::chrono::steady_clock::now();
    loam_counts c = {0, 0, 0, 0, 0};
    int rc = pipe_error(p, j.epoch);
    const bool skip_a = rc != 0;
    if (!rc && j.pre) {
      c = j.counts;  // extracted by loam_pipeline_submit_batch (one launch per kernel for all pipelines of the batch)
      f.c = c;
      memcpy(f.imu, h->imu, sizeof(f.imu));
      const void* src[5] = {h->cur_full, h->cur_sharp, h->cur_less_sharp, h->cur_flat, h->cur_less_flat};
      const int cnt[5] = {c.n_full, c.n_sharp, c.n_less_sharp, c.n_flat, c.n_less_flat};
      cudaStreamWaitEvent(h->st, f.consumed, 0);  // odometry of the sweep that used this slot has finished reading it
      for (int i = 0; i < 5 && !rc; i++) {
        if (f.b[i].ensure((size_t)(cnt[i] + 16) * 16, h->st) != cudaSuccess) rc = LOAM_ECUDA;
        else if (cnt[i]) cudaMemcpyAsync(f.b[i].p, src[i], (size_t)cnt[i] * 16, cudaMemcpyDeviceToDevice, h->st);
      }
      cudaEventRecord(f.ready, h->st);
    } else if (!rc) {
      if (j.slot >= 0) cudaStreamWaitEvent(h->st, p->in_copied[j.slot], 0);
      cudaStreamWaitEvent(h->st, f.consumed, 0);  // odometry of the sweep that used this slot has finished reading it
      DevBuf* ws_out[5] = {&h->sr.full, &h->sr.sharp, &h->sr.less_sharp, &h->sr.flat, &h->sr.less_flat};
      for (int i = 0; i < 5; i++) std::swap(*ws_out[i], f.b[i]);
      g_lg_prof = h->prof.on ? &h->prof : nullptr;
      rc = extract_common(h, j.xyz, j.n, j.stride, nullptr, &c, j.stamp);
      for (int i = 0; i < 5; i++) std::swap(*ws_out[i], f.b[i]);  // the slot keeps the results, the workspace its other set
      if (!rc) {
        f.c = c;
        memcpy(f.imu, h->imu, sizeof(f.imu));
        cudaEventRecord(f.ready, h->st);
      }
    }
    if (j.slot >= 0) p->in_free.release();  // extract_common synchronised: the input slot is free again
    p->busy[0] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t_busy0).count();
    if (rc && !skip_a) pipe_fail(p, rc, j.epoch);
    {
      std::lock_guard<std::mutex> l(p->rm);
      p->partial[j.k].counts = c;
      if (j.pre) p->pre_handled++;  // the copies out of the extraction buffers are enqueued: the next batch may overwrite them
    }
    if (j.pre) p->rcv.notify_all();
    j.slot = fs;
    p->qB.push(j);
  }
}

void stage_b(loam_pipeline* p) {
  cudaSetDevice(p->device);
  loam_handle* h = p->hB;
  for (;;) {
    Job j = p->qB.pop();
    if (j.kind == JOB_RESET) {
      h->lo_inited = false;
      {
        std::lock_guard<std::mutex> l(p->rm);
        p->resets_at_b++;
      }
      p->rcv.notify_all();
    }
    if (j.kind != JOB_SWEEP) {
      p->qC.push(j);
      if (j.kind == JOB_STOP) return;
      continue;
    }
    const auto t_busy0 = std::chrono::steady_clock::now();
    loam_pipeline::Feat& f = p->feat[j.slot];
    loam_odom_result o;
    memset(&o, 0, sizeof(o));
    int rc = pipe_error(p, j.epoch);
    const bool skip_b = rc != 0;
    int ms = -1;
    if (!rc) {
      cudaStreamWaitEvent(h->st, f.ready, 0);
      h->counts = f.c;
      h->cur_full = f.b[0].as<float4>();
      h->cur_sharp = f.b[1].as<float4>();
      h->cur_less_sharp = f.b[2].as<float4>();
      h->cur_flat = f.b[3].as<float4>();
      h->cur_less_flat = f.b[4].as<float4>();
      memcpy(h->imu, f.imu, sizeof(f.imu));
      h->have_features = true;
      rc = loam_odometry_process(h, &o);
      cudaEventRecord(f.consumed, h->st);
    }
    p->feat_free.release();
    p->busy[1] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t_busy0).count();
    if (!rc && o.odom_published && o.fullres_published) {
      p->map_free.acquire();
      ms = (int)(p->map_count++ % PNS);
      loam_pipeline::MapIn& m = p->mapin[ms];
      m.nc = h->n_corner_last; m.ns = h->n_surf_last; m.nf = h->n_fullres3;
      cudaStreamWaitEvent(h->st, m.consumed, 0);
      cudaError_t e = m.corner.ensure((size_t)(m.nc + 16) * 16, h->st);
      if (e == cudaSuccess) e = m.surf.ensure((size_t)(m.ns + 16) * 16, h->st);
      if (e == cudaSuccess) e = m.full.ensure((size_t)(m.nf + 16) * 16, h->st);
      if (e != cudaSuccess) rc = LOAM_ECUDA;
      else {
        if (m.nc) cudaMemcpyAsync(m.corner.p, h->corner_last.p, (size_t)m.nc * 16, cudaMemcpyDeviceToDevice, h->st);
        if (m.ns) cudaMemcpyAsync(m.surf.p, h->surf_last.p, (size_t)m.ns * 16, cudaMemcpyDeviceToDevice, h->st);
        if (m.nf) cudaMemcpyAsync(m.full.p, h->fullres3.p, (size_t)m.nf * 16, cudaMemcpyDeviceToDevice, h->st);
        cudaEventRecord(m.ready, h->st);
      }
    }
    if (rc && !skip_b) pipe_fail(p, rc, j.epoch);
    {
      std::lock_guard<std::mutex> l(p->rm);
      p->partial[j.k].odom = o;
    }
    j.slot = ms;
    j.odom_published = o.odom_published;
    j.full = ms >= 0;
    for (int i = 0; i < 6; i++) j.Tsum[i] = o.transform_sum[i];
    p->qC.push(j);
  }
}

void stage_c(loam_pipeline* p) {
  cudaSetDevice(p->device);
  loam_handle* h = p->hC;
  for (;;) {
    Job j = p->qC.pop();
    if (j.kind == JOB_STOP) {
      p->qD.push(j);
      return;
    }
    if (j.kind != JOB_SWEEP) continue;
    const auto t_busy0 = std::chrono::steady_clock::now();
    loam_map_result mr;
    memset(&mr, 0, sizeof(mr));
    int rc = pipe_error(p, j.epoch);
    const bool skip_c = rc != 0;
    int ran = 0;
    if (!rc && j.odom_published) {
      float hop[6];
      if (h->prm.pose_message_hop) lgh::pose_message_hop(j.Tsum, hop);
      rc = loam_mapping_odometry(h, h->prm.pose_message_hop ? hop : j.Tsum);
    }
    if (j.full) {
      loam_pipeline::MapIn& m = p->mapin[j.slot];
      if (!rc) {
        cudaStreamWaitEvent(h->st, m.ready, 0);
        std::swap(h->corner_last, m.corner);
        std::swap(h->surf_last, m.surf);
        std::swap(h->fullres3, m.full);
        h->n_corner_last = m.nc; h->n_surf_last = m.ns; h->n_fullres3 = m.nf;
        // the run that publishes the surround cloud (LM:1081-1083; the first run after a reset does, LM:434-461) needs the
        // output handle's scratch: wait for the previous surround job only then
        const bool reserve = h->aux && (!h->lm_inited || h->mapFrameCount + 1 >= 5);
        if (reserve) p->aux_free.acquire();
        h->aux_reserved = reserve;
        rc = loam_mapping_process(h, &mr);
        h->aux_reserved = false;
        if (reserve && (rc || mr.n_surround != -1)) p->aux_free.release();  // no surround job was handed over after all
        cudaEventRecord(m.consumed, h->st);
        std::swap(h->corner_last, m.corner);
        std::swap(h->surf_last, m.surf);
        std::swap(h->fullres3, m.full);
        ran = 1;
      }
      p->map_free.release();
    }
    if (rc && !skip_c) pipe_fail(p, rc, j.epoch);
    p->busy[2] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t_busy0).count();
    const bool to_output = !rc && ran && mr.n_surround == -1;  // the surround cloud of this run is still being made
    {
      std::lock_guard<std::mutex> l(p->rm);
      loam_sweep_result& pr = p->partial[j.k];
      pr.map = mr;
      pr.mapping_ran = ran;
      if (!to_output) {
        p->done[j.k] = pr;
        p->partial.erase(j.k);
      }
    }
    if (to_output) {
      j.n = h->aux_ns;
      p->qD.push(j);
    } else {
      p->rcv.notify_all();
    }
  }
}

// Output thread: LM:1092-1094 (VoxelGrid 0.2 m over the gathered surround cubes) on the output handle's stream, then the
// sweep's result is released to loam_pipeline_wait.
void stage_d(loam_pipeline* p) {
  cudaSetDevice(p->device);
  loam_handle* h = p->hD;
  for (;;) {
    Job j = p->qD.pop();
    if (j.kind == JOB_STOP) return;
    int rc = pipe_error(p, j.epoch);
    const bool skip_d = rc != 0;
    int n_out = 0;
    if (!rc) {
      g_lg_prof = h->prof.on ? &h->prof : nullptr;
      rc = (int)h->surround.ensure((size_t)(j.n + 16) * 16, h->st) == (int)cudaSuccess ? LOAM_OK : LOAM_ECUDA;
      if (!rc) {
        std::vector<VoxSegD> segs(1);
        segs[0] = VoxSegD{h->vg_in.as<float4>(), nullptr, h->surround.as<float4>(), nullptr, j.n, 0.2f};
        std::vector<int> cnt;
        rc = voxel_segments(h, segs, cnt);
        if (!rc) {
          n_out = cnt[0];
          h->n_surround = n_out;
          rc = cudaStreamSynchronize(h->st) == cudaSuccess ? LOAM_OK : LOAM_ECUDA;
        }
      }
    }
    p->aux_free.release();
    if (rc && !skip_d) pipe_fail(p, rc, j.epoch);
    {
      std::lock_guard<std::mutex> l(p->rm);
      loam_sweep_result r = p->partial[j.k];
      p->partial.erase(j.k);
      r.map.n_surround = n_out;
      p->done[j.k] = r;
    }
    p->rcv.notify_all();
  }
}

}  // namespace

extern "C" {

int loam_pipeline_create(const loam_params* prm, int device, loam_pipeline** out) {
  if (!out) return LOAM_EINVAL;
  *out = nullptr;
  loam_pipeline* p = new loam_pipeline;
  p->device = device;
  int rc = create_internal(prm, device, 0, &p->hA);
  if (!rc) rc = create_internal(prm, device, 1, &p->hB);
  if (!rc) rc = create_internal(prm, device, 2, &p->hC);
  if (!rc && p->hC->prm.want_surround) rc = create_internal(prm, device, 4, &p->hD);
  if (rc) {
    if (p->hA) loam_destroy(p->hA);
    if (p->hB) loam_destroy(p->hB);
    if (p->hC) loam_destroy(p->hC);
    if (p->hD) loam_destroy(p->hD);
    delete p;
    return rc;
  }
  cudaStreamCreateWithFlags(&p->copy_st, cudaStreamNonBlocking);
  for (int i = 0; i < PNS; i++) {
    cudaEventCreateWithFlags(&p->in_copied[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&p->feat[i].ready, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&p->feat[i].consumed, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&p->mapin[i].ready, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&p->mapin[i].consumed, cudaEventDisableTiming);
  }
  if (p->hD) {
    p->hC->aux = p->hD;
    cudaEventCreateWithFlags(&p->hC->ev_map_done, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&p->hC->ev_aux_read, cudaEventDisableTiming);
  }
  p->tA = std::thread(stage_a, p);
  p->tB = std::thread(stage_b, p);
  p->tC = std::thread(stage_c, p);
  p->tD = std::thread(stage_d, p);
  *out = p;
  return LOAM_OK;
}

loam_handle* loam_pipeline_handle(loam_pipeline* p, int which) {
  if (!p) return nullptr;
  return which == 0 ? p->hA : which == 1 ? p->hB : which == 2 ? p->hC : which == 3 ? p->hD : nullptr;
}

int loam_pipeline_stage_times(loam_pipeline* p, double* out3, int clear) {
  if (!p || !out3) return LOAM_EINVAL;
  for (int i = 0; i < 3; i++) {
    out3[i] = p->busy[i];
    if (clear) p->busy[i] = 0.0;
  }
  return LOAM_OK;
}

int loam_pipeline_destroy(loam_pipeline* p) {
  if (!p) return LOAM_EINVAL;
  Job j;
  memset(&j, 0, sizeof(j));
  j.kind = JOB_STOP;
  p->qA.push(j);
  p->tA.join();
  p->tB.join();
  p->tC.join();
  p->tD.join();
  cudaSetDevice(p->device);
  cudaDeviceSynchronize();
  for (int i = 0; i < PNS; i++) {
    p->in_xyz[i].release();
    for (auto& b : p->feat[i].b) b.release();
    p->mapin[i].corner.release(); p->mapin[i].surf.release(); p->mapin[i].full.release();
    cudaEventDestroy(p->in_copied[i]);
    cudaEventDestroy(p->feat[i].ready); cudaEventDestroy(p->feat[i].consumed);
    cudaEventDestroy(p->mapin[i].ready); cudaEventDestroy(p->mapin[i].consumed);
  }
  cudaStreamDestroy(p->copy_st);
  loam_destroy(p->hA);
  loam_destroy(p->hB);
  loam_destroy(p->hC);
  if (p->hD) loam_destroy(p->hD);
  delete p;
  return LOAM_OK;
}

int loam_pipeline_reset(loam_pipeline* p) {
  if (!p) return LOAM_EINVAL;
  Job j;
  memset(&j, 0, sizeof(j));
  j.kind = JOB_RESET;
  {
    std::lock_guard<std::mutex> l(p->rm);
    p->epoch++;  // an error of the epoch that ends here no longer blocks submit / stages
    p->resets_pushed++;
  }
  p->qA.push(j);
  return LOAM_OK;
}
const char* loam_pipeline_last_error(loam_pipeline* p) {
  static thread_local char buf[512];
  if (!p) return "";
  std::lock_guard<std::mutex> l(p->rm);
  memcpy(buf, p->err_text, sizeof(buf));
  return buf;
}

static int pipeline_submit(loam_pipeline* p, const float* xyz, int n, int stride_bytes, bool host, double stamp) {
  // any point_step >= 12 and any alignment is accepted, like loam_extract: extract_common repacks unaligned layouts
  if (!p || n < 0 || (!xyz && n > 0) || stride_bytes < 12) return LOAM_EINVAL;
  long long epoch;
  {
    std::lock_guard<std::mutex> l(p->rm);
    epoch = p->epoch;
  }
  if (int e = pipe_error(p, epoch)) return e;
  LG_CHECK(cudaSetDevice(p->device));
  Job j;
  memset(&j, 0, sizeof(j));
  j.kind = JOB_SWEEP;
  j.n = n;
  j.stride = stride_bytes;
  j.slot = -1;
  j.xyz = xyz;
  j.epoch = epoch;
  j.stamp = stamp;
  if (host) {
    p->in_free.acquire();
    const int s = (int)(p->in_count % PNS);  // the slot is only taken once the copy has succeeded
    cudaError_t e = p->in_xyz[s].ensure((size_t)n * stride_bytes + 64, p->copy_st);
    if (e == cudaSuccess && n) e = cudaMemcpyAsync(p->in_xyz[s].p, xyz, (size_t)n * stride_bytes, cudaMemcpyHostToDevice, p->copy_st);
    if (e == cudaSuccess) e = cudaEventRecord(p->in_copied[s], p->copy_st);
    if (e == cudaSuccess) e = cudaEventSynchronize(p->in_copied[s]);  // the caller may reuse its buffer as soon as we return
    if (e != cudaSuccess) {
      lg_set_error(cudaGetErrorString(e), __FILE__, __LINE__);
      {
        std::lock_guard<std::mutex> l(p->rm);
        snprintf(p->err_text, sizeof(p->err_text), "%s", g_cuda_err);
      }
      p->in_free.release();
      return LOAM_ECUDA;
    }
    p->in_count++;
    p->hA->h2d_bytes += (long long)n * stride_bytes;
    j.slot = s;
    j.xyz = p->in_xyz[s].as<float>();
  }
  {
    std::lock_guard<std::mutex> l(p->rm);
    j.k = p->next_submit++;
    memset(&p->partial[j.k], 0, sizeof(loam_sweep_result));
    p->epoch_of[j.k] = epoch;
  }
  p->qA.push(j);
  return LOAM_OK;
}
int loam_pipeline_submit(loam_pipeline* p, const float* xyz_host, int n, int stride_bytes, double stamp) {
  return pipeline_submit(p, xyz_host, n, stride_bytes, true, stamp);
}
int loam_pipeline_submit_device(loam_pipeline* p, const float* xyz_dev, int n, int stride_bytes, double stamp) {
  return pipeline_submit(p, xyz_dev, n, stride_bytes, false, stamp);
}
// One /imu/data message for the pipeline's extraction stage (loam_imu_push), applied in submission order: sweeps submitted
// before it do not see it, sweeps submitted after it do -- the pipelined results equal the blocking calls'.
int loam_pipeline_imu_push(loam_pipeline* p, double stamp, const double* orientation_xyzw, const double* angular_velocity,
                           const double* linear_acceleration) {
  if (!p || !orientation_xyzw || !angular_velocity || !linear_acceleration) return LOAM_EINVAL;
  Job j;
  memset(&j, 0, sizeof(j));
  j.kind = JOB_IMU;
  j.stamp = stamp;
  for (int i = 0; i < 4; i++) j.imu_msg[i] = orientation_xyzw[i];
  for (int i = 0; i < 3; i++) j.imu_msg[4 + i] = angular_velocity[i], j.imu_msg[7 + i] = linear_acceleration[i];
  {
    std::lock_guard<std::mutex> l(p->rm);
    p->imu_pushed++;
  }
  p->qA.push(j);
  return LOAM_OK;
}

// One sweep for each of B pipelines (independent sequences on one device): the extraction of all B sweeps is done here, in
// the caller's thread, with loam_extract_batch (one launch per kernel for the whole batch); every pipeline's stage A then
// only hands the clouds to its odometry stage.  Results per pipeline are those of loam_pipeline_submit.  Blocks for the
// extraction (~0.3 ms for eight VLP-16 sweeps); do not mix with loam_pipeline_submit calls in flight on the same pipelines.
int loam_pipeline_submit_batch(loam_pipeline* const* ps, int B, const float* const* xyz_host, const int* n, int stride_bytes, const double* stamps) {
  if (!ps || B < 1 || B > 256 || !xyz_host || !n) return LOAM_EINVAL;
  std::vector<loam_handle*> hs(B);
  std::vector<long long> epochs(B);
  for (int b = 0; b < B; b++) {
    if (!ps[b] || ps[b]->device != ps[0]->device) return LOAM_EINVAL;
    hs[b] = ps[b]->hA;
    std::unique_lock<std::mutex> l(ps[b]->rm);
    epochs[b] = ps[b]->epoch;
    // stage A must have handed on the previous batch's clouds before they are overwritten
    ps[b]->rcv.wait(l, [&] { return ps[b]->pre_handled == ps[b]->pre_submitted && ps[b]->imu_handled == ps[b]->imu_pushed; });
  }
  for (int b = 0; b < B; b++)
    if (int e = pipe_error(ps[b], epochs[b])) return e;
  std::vector<loam_counts> counts(B);
  int rc = loam_extract_batch(hs.data(), B, xyz_host, n, stride_bytes, stamps, counts.data());
  if (rc) return rc;
  for (int b = 0; b < B; b++) {
    loam_pipeline* p = ps[b];
    Job j;
    memset(&j, 0, sizeof(j));
    j.kind = JOB_SWEEP;
    j.n = n[b];
    j.stride = stride_bytes;
    j.slot = -1;
    j.epoch = epochs[b];
    j.pre = 1;
    j.counts = counts[b];
    {
      std::lock_guard<std::mutex> l(p->rm);
      j.k = p->next_submit++;
      memset(&p->partial[j.k], 0, sizeof(loam_sweep_result));
      p->epoch_of[j.k] = epochs[b];
      p->pre_submitted++;
    }
    p->qA.push(j);
  }
  return LOAM_OK;
}

// One sweep for each of B pipelines with extraction AND scan-to-scan odometry batched in lock-step in the caller's thread
// (loam_extract_batch + loam_odometry_process_batch: one launch per kernel and round for all B sequences); the mapping and
// output stages stay per pipeline.  Per-pipeline results are those of loam_pipeline_submit.  Blocks for the extraction and
// the odometry of the batch (and for a free mapping slot); use it INSTEAD of loam_pipeline_submit on these pipelines.
int loam_pipeline_submit_lockstep(loam_pipeline* const* ps, int B, const float* const* xyz_host, const int* n, int stride_bytes, const double* stamps) {
  if (!ps || B < 1 || B > 256 || !xyz_host || !n) return LOAM_EINVAL;
  std::vector<loam_handle*> hA(B), hB(B);
  std::vector<long long> epochs(B);
  for (int b = 0; b < B; b++) {
    if (!ps[b] || ps[b]->device != ps[0]->device) return LOAM_EINVAL;
    hA[b] = ps[b]->hA;
    hB[b] = ps[b]->hB;
    std::unique_lock<std::mutex> l(ps[b]->rm);
    // a reset travels through the stage threads: the odometry handle is only touched here once it has arrived
    ps[b]->rcv.wait(l, [&] {
      return ps[b]->resets_at_b == ps[b]->resets_pushed && ps[b]->pre_handled == ps[b]->pre_submitted && ps[b]->imu_handled == ps[b]->imu_pushed;
    });
    epochs[b] = ps[b]->epoch;
  }
  for (int b = 0; b < B; b++)
    if (int e = pipe_error(ps[b], epochs[b])) return e;
  std::vector<loam_counts> counts(B);
  int rc = loam_extract_batch(hA.data(), B, xyz_host, n, stride_bytes, stamps, counts.data());
  if (rc) return rc;
  for (int b = 0; b < B; b++) {  // the odometry handles read the features where the extraction left them
    loam_handle *a = hA[b], *h = hB[b];
    h->counts = counts[b];
    h->cur_full = a->cur_full; h->cur_sharp = a->cur_sharp; h->cur_less_sharp = a->cur_less_sharp;
    h->cur_flat = a->cur_flat; h->cur_less_flat = a->cur_less_flat;
    for (int i = 0; i < 12; i++) h->imu[i] = a->imu[i];
    h->have_features = true;
  }
  std::vector<loam_odom_result> od(B);
  rc = loam_odometry_process_batch(hB.data(), B, od.data());
  if (rc) return rc;
  for (int b = 0; b < B; b++) {  // what the odometry stage does after its sweep: hand the clouds to the mapping stage
    loam_pipeline* p = ps[b];
    loam_handle* h = hB[b];
    const loam_odom_result& o = od[b];
    int ms = -1;
    if (o.odom_published && o.fullres_published) {
      p->map_free.acquire();
      ms = (int)(p->map_count++ % PNS);
      loam_pipeline::MapIn& m = p->mapin[ms];
      m.nc = h->n_corner_last; m.ns = h->n_surf_last; m.nf = h->n_fullres3;
      cudaStreamWaitEvent(h->st, m.consumed, 0);
      cudaError_t e = m.corner.ensure((size_t)(m.nc + 16) * 16, h->st);
      if (e == cudaSuccess) e = m.surf.ensure((size_t)(m.ns + 16) * 16, h->st);
      if (e == cudaSuccess) e = m.full.ensure((size_t)(m.nf + 16) * 16, h->st);
      if (e != cudaSuccess) {
        lg_set_error(cudaGetErrorString(e), __FILE__, __LINE__);
        p->map_free.release();
        return LOAM_ECUDA;
      }
      if (m.nc) cudaMemcpyAsync(m.corner.p, h->corner_last.p, (size_t)m.nc * 16, cudaMemcpyDeviceToDevice, h->st);
      if (m.ns) cudaMemcpyAsync(m.surf.p, h->surf_last.p, (size_t)m.ns * 16, cudaMemcpyDeviceToDevice, h->st);
      if (m.nf) cudaMemcpyAsync(m.full.p, h->fullres3.p, (size_t)m.nf * 16, cudaMemcpyDeviceToDevice, h->st);
      cudaEventRecord(m.ready, h->st);
    }
    Job j;
    memset(&j, 0, sizeof(j));
    j.kind = JOB_SWEEP;
    j.n = n[b];
    j.stride = stride_bytes;
    j.epoch = epochs[b];
    j.slot = ms;
    j.odom_published = o.odom_published;
    j.full = ms >= 0;
    for (int i = 0; i < 6; i++) j.Tsum[i] = o.transform_sum[i];
    {
      std::lock_guard<std::mutex> l(p->rm);
      j.k = p->next_submit++;
      loam_sweep_result& pr = p->partial[j.k];
      memset(&pr, 0, sizeof(pr));
      pr.counts = counts[b];
      pr.odom = o;
      p->epoch_of[j.k] = epochs[b];
    }
    p->qC.push(j);
  }
  return LOAM_OK;
}

int loam_pipeline_wait(loam_pipeline* p, loam_sweep_result* out) {
  if (!p || !out) return LOAM_EINVAL;
  std::unique_lock<std::mutex> l(p->rm);
  if (p->next_wait >= p->next_submit) return LOAM_ESTATE;
  const long long k = p->next_wait;
  p->rcv.wait(l, [&] { return p->done.count(k) > 0; });
  *out = p->done[k];
  p->done.erase(k);
  p->next_wait++;
  const long long ep = p->epoch_of[k];
  p->epoch_of.erase(k);
  return p->error_epoch.load() == ep ? p->error.load() : LOAM_OK;  // an error is reported for the sweeps of its epoch only
}

void* loam_pipeline_stream(loam_pipeline* p, int which) {
  if (!p || which < 0 || which > 2) return nullptr;
  loam_handle* hs[3] = {p->hA, p->hB, p->hC};
  return (void*)hs[which]->st;
}

int loam_pipeline_pending(loam_pipeline* p) {
  if (!p) return LOAM_EINVAL;
  std::lock_guard<std::mutex> l(p->rm);
  return (int)(p->next_submit - p->next_wait);
}

int loam_pipeline_stats(loam_pipeline* p, long long out4[4]) {
  if (!p || !out4) return LOAM_EINVAL;
  for (int i = 0; i < 4; i++) out4[i] = 0;
  for (loam_handle* h : {p->hA, p->hB, p->hC, p->hD}) {
    if (!h) continue;
    out4[0] += h->launches; out4[1] += h->h2d_bytes; out4[2] += h->d2h_bytes; out4[3] += h->syncs;
  }
  return LOAM_OK;
}

}  // extern "C"
