// Host-side check of gpscalibration_b200/csrc/lg_libm.cuh (the device versions of sinf/cosf/atanf/atan2f) against the
// host libm, bit for bit.  Usage: libm_port_check <millions of samples>.  Prints the four mismatch counts.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "lg_libm.cuh"

static uint64_t s = 88172645463325252ull;
static inline uint64_t rnd(void) {
  s ^= s << 13;
  s ^= s >> 7;
  s ^= s << 17;
  return s;
}
static inline float urand(float lo, float hi) { return lo + (hi - lo) * (float)((rnd() >> 40) * (1.0 / 16777216.0)); }

int main(int argc, char** argv) {
  long n = (argc > 1 ? atol(argv[1]) : 4) * 1000000L, bs = 0, bc = 0, ba = 0, ba2 = 0;
  for (long i = 0; i < n; i++) {
    float r = (i & 3) == 0 ? urand(-0.2f, 0.2f) : (i & 3) == 1 ? urand(-0.8f, 0.8f) : (i & 3) == 2 ? urand(-7.f, 7.f) : urand(-100.f, 100.f);
    if (i < 4096) r = ldexpf(urand(-1.f, 1.f), -(int)(i % 40));  // tiny arguments
    bs += lgm_asuint(sinf(r)) != lgm_asuint(lgm_sinf(r));
    bc += lgm_asuint(cosf(r)) != lgm_asuint(lgm_cosf(r));
    float a = (i & 1) ? urand(-3.f, 3.f) : urand(-0.5f, 0.5f);
    ba += lgm_asuint(atanf(a)) != lgm_asuint(lgm_atanf(a));
    float y = urand(-100.f, 100.f), x = urand(-100.f, 100.f);
    if ((i & 1023) == 0) y = 0.f;
    if ((i & 1023) == 1) x = 0.f;
    if ((i & 1023) == 2) x = 1.f;
    ba2 += lgm_asuint(atan2f(y, x)) != lgm_asuint(lgm_atan2f(y, x));
  }
  printf("%ld %ld %ld %ld\n", bs, bc, ba, ba2);
  return 0;
}
