// C ABI of the synthetic sweep generator (see synth.h).  Built into libloamsynth.so (host only, no CUDA).
#include "synth.h"

#include <map>
#include <memory>
#include <mutex>

using namespace loamsynth;

namespace {
struct Ctx {
  Scene scene;
  SensorModel sensor;
  uint64_t seed;
};
}  // namespace

extern "C" {

// sensor_kind: 0 reference-ring-table 16x1800, 1 true VLP-16 angles 16x1800, 2 HDL-64-shaped 64x1875.
// scene_kind: 0 ring road, 1 ring road + dense blocks.
void* loamsynth_create(int sensor_kind, int scene_kind, unsigned long long seed) {
  Ctx* c = new Ctx;
  c->scene = make_scene(scene_kind, 0xC0FFEEull);  // the scene is fixed; `seed` only drives the range noise
  c->sensor = make_sensor(sensor_kind);
  c->seed = seed;
  return c;
}
void loamsynth_destroy(void* h) { delete (Ctx*)h; }
int loamsynth_max_points(void* h) {
  Ctx* c = (Ctx*)h;
  return c->sensor.n_rings * c->sensor.n_cols;
}
int loamsynth_rings(void* h) { return ((Ctx*)h)->sensor.n_rings; }
// Writes up to max_points xyz triples; returns the point count.  pose6 (optional) = x,y,z,yaw,pitch,roll at sweep start.
int loamsynth_sweep(void* h, int sweep_id, double t_offset, float* xyz, double* pose6) {
  Ctx* c = (Ctx*)h;
  if (pose6) {
    Pose p = trajectory(t_offset + sweep_id);
    pose6[0] = p.x; pose6[1] = p.y; pose6[2] = p.z; pose6[3] = p.yaw; pose6[4] = p.pitch; pose6[5] = p.roll;
  }
  return synth_sweep(c->scene, c->sensor, c->seed + (uint64_t)sweep_id, sweep_id, t_offset, xyz);
}

}  // extern "C"
