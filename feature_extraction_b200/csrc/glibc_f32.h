// glibc_f32.h — atan2f and acosf exactly as glibc <= 2.40 computes them.
//
// pcl::ShapeContext3DEstimation::computePoint (PCL 1.8.0 features/impl/3dsc.hpp; called from the
// reference at src/feature_extraction_node.cpp:353) bins every neighbour by
//     phi   = rad2deg(atan2(cross.norm(), x_axis.dot(proj)))     float arguments -> atan2f
//     theta = rad2deg(acosf(clamp(normal.dot(no), -1, 1)))
// so the bin a neighbour lands in depends on the last bit of libm's float atan2/acos.  Until the
// CORE-MATH rewrite in glibc 2.41 those are the Sun fdlibm float routines
// (sysdeps/ieee754/flt-32/e_atan2f.c, s_atanf.c, e_acosf.c; no x86-64 multiarch/FMA variants: the
// symbols are plain functions in libm.so.6), < 1 ulp but not correctly rounded.  This header restates the
// published fdlibm algorithm operation by operation in unfused float arithmetic, so that the device
// rounds the way the host's libm does.  Checked against the libm of this image (glibc 2.39):
// acosf over every float in [-1, 1], atanf over every float, atan2f over 10^9 argument pairs
// (tests/test_glibc_f32.py); and on the device against the same libm (tests/test_gpu_libm.py).
//
// Usable from host code (plain C++, build with -ffp-contract=off) and from device code.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define FE_HD __host__ __device__ __forceinline__
#else
#define FE_HD inline
#endif

namespace fe {
namespace glibc {

// every operation individually rounded to float: no contraction on either side
FE_HD float f_mul(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fmul_rn(a, b);
#else
  return a * b;
#endif
}
FE_HD float f_add(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fadd_rn(a, b);
#else
  return a + b;
#endif
}
FE_HD float f_sub(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fsub_rn(a, b);
#else
  return a - b;
#endif
}
FE_HD float f_div(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fdiv_rn(a, b);
#else
  return a / b;
#endif
}
FE_HD float f_sqrt(float a) {
#if defined(__CUDA_ARCH__)
  return __fsqrt_rn(a);
#else
  return __builtin_sqrtf(a);
#endif
}
FE_HD int32_t f_bits(float a) {
#if defined(__CUDA_ARCH__)
  return __float_as_int(a);
#else
  int32_t i;
  memcpy(&i, &a, 4);
  return i;
#endif
}
FE_HD float f_from_bits(int32_t i) {
#if defined(__CUDA_ARCH__)
  return __int_as_float(i);
#else
  float a;
  memcpy(&a, &i, 4);
  return a;
#endif
}

// s_atanf.c: argument reduction to |x| < 7/16 around 0.5, 1, 1.5, inf; odd/even split polynomial
FE_HD float atanf_fdlibm(float x) {
  const float aT0 = 3.3333334327e-01f, aT1 = -2.0000000298e-01f, aT2 = 1.4285714924e-01f, aT3 = -1.1111110449e-01f,
              aT4 = 9.0908870101e-02f, aT5 = -7.6918758452e-02f, aT6 = 6.6610731184e-02f, aT7 = -5.8335702866e-02f,
              aT8 = 4.9768779427e-02f, aT9 = -3.6531571299e-02f, aT10 = 1.6285819933e-02f;
  const int32_t hx = f_bits(x);
  const int32_t ix = hx & 0x7fffffff;
  int id;
  float hi = 0.0f, lo = 0.0f;
  if (ix >= 0x4c000000) {  // |x| >= 2^25
    if (ix > 0x7f800000) return f_add(x, x);  // NaN
    const float r = f_add(1.5707962513e+00f, 7.5497894159e-08f);
    return (hx > 0) ? r : -r;
  }
  if (ix < 0x3ee00000) {  // |x| < 0.4375
    if (ix < 0x31000000) return x;  // |x| < 2^-29
    id = -1;
  } else {
    x = f_from_bits(ix);  // fabsf
    if (ix < 0x3f980000) {    // |x| < 1.1875
      if (ix < 0x3f300000) {  // 7/16 <= |x| < 11/16
        id = 0; hi = 4.6364760399e-01f; lo = 5.0121582440e-09f;
        x = f_div(f_sub(f_mul(2.0f, x), 1.0f), f_add(2.0f, x));
      } else {                // 11/16 <= |x| < 19/16
        id = 1; hi = 7.8539812565e-01f; lo = 3.7748947079e-08f;
        x = f_div(f_sub(x, 1.0f), f_add(x, 1.0f));
      }
    } else {
      if (ix < 0x401c0000) {  // |x| < 2.4375
        id = 2; hi = 9.8279368877e-01f; lo = 3.4473217170e-08f;
        x = f_div(f_sub(x, 1.5f), f_add(1.0f, f_mul(1.5f, x)));
      } else {                // 2.4375 <= |x| < 2^25
        id = 3; hi = 1.5707962513e+00f; lo = 7.5497894159e-08f;
        x = f_div(-1.0f, x);
      }
    }
  }
  const float z = f_mul(x, x);
  const float w = f_mul(z, z);
  float s1 = f_add(aT8, f_mul(w, aT10));
  s1 = f_add(aT6, f_mul(w, s1));
  s1 = f_add(aT4, f_mul(w, s1));
  s1 = f_add(aT2, f_mul(w, s1));
  s1 = f_add(aT0, f_mul(w, s1));
  s1 = f_mul(z, s1);
  float s2 = f_add(aT7, f_mul(w, aT9));
  s2 = f_add(aT5, f_mul(w, s2));
  s2 = f_add(aT3, f_mul(w, s2));
  s2 = f_add(aT1, f_mul(w, s2));
  s2 = f_mul(w, s2);
  if (id < 0) return f_sub(x, f_mul(x, f_add(s1, s2)));
  const float r = f_sub(hi, f_sub(f_sub(f_mul(x, f_add(s1, s2)), lo), x));
  return (hx < 0) ? -r : r;
}

// e_atan2f.c
FE_HD float atan2f_fdlibm(float y, float x) {
  const float tiny = 1.0e-30f, pi_o_4 = 7.8539818525e-01f, pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f,
              pi_lo = -8.7422776573e-08f;
  const int32_t hx = f_bits(x), hy = f_bits(y);
  const int32_t ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
  if (ix > 0x7f800000 || iy > 0x7f800000) return f_add(x, y);  // NaN
  if (hx == 0x3f800000) return atanf_fdlibm(y);                 // x == 1
  const int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);            // 2*sign(x) + sign(y)
  if (iy == 0) {
    switch (m) {
      case 0:
      case 1: return y;
      case 2: return f_add(pi, tiny);
      default: return f_sub(-pi, tiny);
    }
  }
  if (ix == 0) return (hy < 0) ? f_sub(-pi_o_2, tiny) : f_add(pi_o_2, tiny);
  if (ix == 0x7f800000) {
    if (iy == 0x7f800000) {
      switch (m) {
        case 0: return f_add(pi_o_4, tiny);
        case 1: return f_sub(-pi_o_4, tiny);
        case 2: return f_add(f_mul(3.0f, pi_o_4), tiny);
        default: return f_sub(f_mul(-3.0f, pi_o_4), tiny);
      }
    } else {
      switch (m) {
        case 0: return 0.0f;
        case 1: return -0.0f;
        case 2: return f_add(pi, tiny);
        default: return f_sub(-pi, tiny);
      }
    }
  }
  if (iy == 0x7f800000) return (hy < 0) ? f_sub(-pi_o_2, tiny) : f_add(pi_o_2, tiny);
  const int32_t k = (iy - ix) >> 23;
  float z;
  if (k > 60) z = f_add(pi_o_2, f_mul(0.5f, pi_lo));  // |y/x| > 2^60
  else if (hx < 0 && k < -60) z = 0.0f;               // |y|/x < -2^60
  else z = atanf_fdlibm(f_from_bits(f_bits(f_div(y, x)) & 0x7fffffff));
  switch (m) {
    case 0: return z;
    case 1: return f_from_bits(f_bits(z) ^ (int32_t)0x80000000);
    case 2: return f_sub(pi, f_sub(z, pi_lo));
    default: return f_sub(f_sub(z, pi_lo), pi);
  }
}

// e_acosf.c
FE_HD float acosf_fdlibm(float x) {
  const float pi = 3.1415925026e+00f, pio2_hi = 1.5707962513e+00f, pio2_lo = 7.5497894159e-08f;
  const float pS0 = 1.6666667163e-01f, pS1 = -3.2556581497e-01f, pS2 = 2.0121252537e-01f, pS3 = -4.0055535734e-02f,
              pS4 = 7.9153501429e-04f, pS5 = 3.4793309169e-05f;
  const float qS1 = -2.4033949375e+00f, qS2 = 2.0209457874e+00f, qS3 = -6.8828397989e-01f, qS4 = 7.7038154006e-02f;
  const int32_t hx = f_bits(x);
  const int32_t ix = hx & 0x7fffffff;
  if (ix == 0x3f800000) {  // |x| == 1
    if (hx > 0) return 0.0f;
    return f_add(pi, f_mul(2.0f, pio2_lo));
  }
  if (ix > 0x3f800000) return f_div(f_sub(x, x), f_sub(x, x));  // |x| > 1 or NaN -> NaN
  float z;
  if (ix < 0x3f000000) {  // |x| < 0.5
    if (ix <= 0x32800000) return f_add(pio2_hi, pio2_lo);  // |x| < 2^-26
    z = f_mul(x, x);
  } else if (hx < 0) {
    z = f_mul(f_add(1.0f, x), 0.5f);
  } else {
    z = f_mul(f_sub(1.0f, x), 0.5f);
  }
  float p = f_add(pS4, f_mul(z, pS5));
  p = f_add(pS3, f_mul(z, p));
  p = f_add(pS2, f_mul(z, p));
  p = f_add(pS1, f_mul(z, p));
  p = f_add(pS0, f_mul(z, p));
  p = f_mul(z, p);
  float q = f_add(qS3, f_mul(z, qS4));
  q = f_add(qS2, f_mul(z, q));
  q = f_add(qS1, f_mul(z, q));
  q = f_add(1.0f, f_mul(z, q));
  const float r = f_div(p, q);
  if (ix < 0x3f000000) return f_sub(pio2_hi, f_sub(x, f_sub(pio2_lo, f_mul(r, x))));
  const float s = f_sqrt(z);
  if (hx < 0) {
    const float w = f_sub(f_mul(r, s), pio2_lo);
    return f_sub(pi, f_mul(2.0f, f_add(s, w)));
  }
  const float df = f_from_bits(f_bits(s) & (int32_t)0xfffff000);
  const float c = f_div(f_sub(z, f_mul(df, df)), f_add(s, df));
  const float w = f_add(f_mul(r, s), c);
  return f_mul(2.0f, f_add(df, w));
}

}  // namespace glibc
}  // namespace fe
