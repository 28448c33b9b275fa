// Bit-faithful device versions of the four libm binary32 functions the reference's hot loops call per point:
//   sin(float) / cos(float)   TransformToStart / TransformToEnd, LO:138-148, 170-198   (-> glibc sinf / cosf)
//   atan(float)               ring angle,        SR:297                               (-> glibc atanf)
//   atan2(float, float)       azimuth,           SR:267-270, 340                      (-> glibc atan2f)
// The reference's results (and therefore the oracle's) are whatever the host libm returns; CUDA's own sinf/atan2f
// differ from it by an ulp now and then, which is enough to flip a nearest-neighbour tie or a convergence test
// twenty iterations later.  These are re-statements of the published algorithms glibc 2.39 (this image) uses:
//   sinf/cosf : ARM "optimized routines" sincosf — fp64 range reduction + fp64 minimax polynomials, one rounding.
//               x86-64 glibc dispatches to its FMA build (sysdeps/x86_64/fpu/multiarch/s_sinf-fma.c), so every
//               a + b*c below is an explicit fused multiply-add.
//   atanf/atan2f : Sun fdlibm (s_atanf.c / e_atan2f.c), plain fp32, no fusion.
// tests/test_libm_port.py compiles this header for the host and checks it bit-for-bit against the host libm on tens of
// millions of arguments (all ranges the hot path produces), so the oracle can keep calling libm like the reference.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define LGM_HD __host__ __device__ __forceinline__
#else
#define LGM_HD static inline
#endif

LGM_HD uint32_t lgm_asuint(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
}
LGM_HD float lgm_asfloat(uint32_t u) {
  float f;
  memcpy(&f, &u, 4);
  return f;
}
LGM_HD uint32_t lgm_abstop12(float x) { return (lgm_asuint(x) >> 20) & 0x7ff; }

// polynomial / reduction constants of __sincosf_table[0]; table[1] only flips the signs of c0..c4
#define LGM_HPI_INV 0x1.45F306DC9C883p+23 /* 2/pi * 2^24 */
#define LGM_HPI 0x1.921FB54442D18p0
#define LGM_C0 0x1p0
#define LGM_C1 -0x1.ffffffd0c621cp-2
#define LGM_C2 0x1.55553e1068f19p-5
#define LGM_C3 -0x1.6c087e89a359dp-10
#define LGM_C4 0x1.99343027bf8c3p-16
#define LGM_S1 -0x1.555545995a603p-3
#define LGM_S2 0x1.1107605230bc4p-7
#define LGM_S3 -0x1.994eb3774cf24p-13

// sinf_poly: n even -> sine polynomial, n odd -> cosine polynomial (flip = table[1], i.e. negated cosine coefficients)
LGM_HD float lgm_sincos_poly(double x, double x2, int n, bool flip) {
  if ((n & 1) == 0) {
    double x3 = x * x2;
    double s1 = fma(x2, LGM_S3, LGM_S2);
    double x7 = x3 * x2;
    double s = fma(x3, LGM_S1, x);
    return (float)fma(x7, s1, s);
  } else {
    const double sg = flip ? -1.0 : 1.0;
    double x4 = x2 * x2;
    double c2 = fma(x2, sg * LGM_C4, sg * LGM_C3);
    double c1 = fma(x2, sg * LGM_C1, sg * LGM_C0);
    double x6 = x4 * x2;
    double c = fma(x4, sg * LGM_C2, c1);
    return (float)fma(x6, c2, c);
  }
}

LGM_HD double lgm_reduce_fast(double x, int* np) {
  double r = x * LGM_HPI_INV;
  int n = ((int32_t)r + 0x800000) >> 24;
  *np = n;
  return fma(-(double)n, LGM_HPI, x);
}

// Valid for |y| < 120 (the hot path's angles are a few radians at most); beyond that fall back to fp64 evaluation.
LGM_HD float lgm_sinf(float y) {
  double x = y;
  if (lgm_abstop12(y) < lgm_abstop12(0x1.921FB6p-1f)) {  // |y| < pi/4
    double s = x * x;
    if (lgm_abstop12(y) < lgm_abstop12(0x1p-12f)) return y;
    return lgm_sincos_poly(x, s, 0, false);
  } else if (lgm_abstop12(y) < lgm_abstop12(120.0f)) {
    int n;
    x = lgm_reduce_fast(x, &n);
    double s = (((n & 3) == 1) || ((n & 3) == 2)) ? -1.0 : 1.0;  // sign[] = {1, -1, -1, 1}
    return lgm_sincos_poly(x * s, x * x, n, (n & 2) != 0);
  }
  return (float)sin(x);
}

LGM_HD float lgm_cosf(float y) {
  double x = y;
  if (lgm_abstop12(y) < lgm_abstop12(0x1.921FB6p-1f)) {
    double x2 = x * x;
    if (lgm_abstop12(y) < lgm_abstop12(0x1p-12f)) return 1.0f;
    return lgm_sincos_poly(x, x2, 1, false);
  } else if (lgm_abstop12(y) < lgm_abstop12(120.0f)) {
    int n;
    x = lgm_reduce_fast(x, &n);
    double s = (((n & 3) == 1) || ((n & 3) == 2)) ? -1.0 : 1.0;
    return lgm_sincos_poly(x * s, x * x, n ^ 1, (n & 2) != 0);
  }
  return (float)cos(x);
}

// ---- fdlibm s_atanf.c
LGM_HD float lgm_atanf(float x) {
  const float atanhi[4] = {4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f, 1.5707962513e+00f};
  const float atanlo[4] = {5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f, 7.5497894159e-08f};
  const float aT[11] = {3.3333334327e-01f, -2.0000000298e-01f, 1.4285714924e-01f, -1.1111110449e-01f, 9.0908870101e-02f, -7.6918758452e-02f,
                        6.6610731184e-02f, -5.8335702866e-02f, 4.9768779427e-02f, -3.6531571299e-02f, 1.6285819933e-02f};
  const float one = 1.0f;
  float w, s1, s2, z;
  int32_t hx = (int32_t)lgm_asuint(x);
  int32_t ix = hx & 0x7fffffff;
  int id;
  if (ix >= 0x4c000000) {  // |x| >= 2^25
    if (ix > 0x7f800000) return x + x;
    if (hx > 0) return atanhi[3] + atanlo[3];
    return -atanhi[3] - atanlo[3];
  }
  if (ix < 0x3ee00000) {  // |x| < 0.4375
    if (ix < 0x31000000) return x;  // |x| < 2^-29
    id = -1;
  } else {
    x = fabsf(x);
    if (ix < 0x3f980000) {    // |x| < 1.1875
      if (ix < 0x3f300000) {  // 7/16 <= |x| < 11/16
        id = 0;
        x = (2.0f * x - one) / (2.0f + x);
      } else {  // 11/16 <= |x| < 19/16
        id = 1;
        x = (x - one) / (x + one);
      }
    } else {
      if (ix < 0x401c0000) {  // |x| < 2.4375
        id = 2;
        x = (x - 1.5f) / (one + 1.5f * x);
      } else {  // 2.4375 <= |x| < 2^66
        id = 3;
        x = -1.0f / x;
      }
    }
  }
  z = x * x;
  w = z * z;
  s1 = z * (aT[0] + w * (aT[2] + w * (aT[4] + w * (aT[6] + w * (aT[8] + w * aT[10])))));
  s2 = w * (aT[1] + w * (aT[3] + w * (aT[5] + w * (aT[7] + w * aT[9]))));
  if (id < 0) return x - x * (s1 + s2);
  z = atanhi[id] - ((x * (s1 + s2) - atanlo[id]) - x);
  return (hx < 0) ? -z : z;
}

// ---- fdlibm e_atan2f.c
LGM_HD float lgm_atan2f(float y, float x) {
  const float tiny = 1.0e-30f, pi_o_4 = 7.8539818525e-01f, pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f, pi_lo = -8.7422776573e-08f;
  float z;
  int32_t hx = (int32_t)lgm_asuint(x), hy = (int32_t)lgm_asuint(y);
  int32_t ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
  if (ix > 0x7f800000 || iy > 0x7f800000) return x + y;
  if (hx == 0x3f800000) return lgm_atanf(y);
  int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
  if (iy == 0) {
    switch (m) {
      case 0:
      case 1: return y;
      case 2: return pi + tiny;
      default: return -pi - tiny;
    }
  }
  if (ix == 0) return (hy < 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;
  if (ix == 0x7f800000) {
    if (iy == 0x7f800000) {
      switch (m) {
        case 0: return pi_o_4 + tiny;
        case 1: return -pi_o_4 - tiny;
        case 2: return 3.0f * pi_o_4 + tiny;
        default: return -3.0f * pi_o_4 - tiny;
      }
    } else {
      switch (m) {
        case 0: return 0.0f;
        case 1: return -0.0f;
        case 2: return pi + tiny;
        default: return -pi - tiny;
      }
    }
  }
  if (iy == 0x7f800000) return (hy < 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;
  int32_t k = (iy - ix) >> 23;
  if (k > 60) z = pi_o_2 + 0.5f * pi_lo;
  else if (hx < 0 && k < -60) z = 0.0f;
  else z = lgm_atanf(fabsf(y / x));
  switch (m) {
    case 0: return z;
    case 1: return lgm_asfloat(lgm_asuint(z) ^ 0x80000000u);
    case 2: return pi - (z - pi_lo);
    default: return (z - pi_lo) - pi;
  }
}
