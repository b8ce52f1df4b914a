// atanf / atan2f with the results of the fdlibm float routines (Sun Microsystems' freely distributable libm; float conversion by
// I. L. Taylor) — the algorithm glibc <= 2.40 ships for x86-64 (sysdeps/ieee754/flt-32/{s_atanf,e_atan2f}.c).  CUDA's own atanf / atan2f
// are 1-2 ulp routines with different roundings, and the range-image pixel a point falls into (image_segmenter.hpp:103,119) depends on
// the last bit; restating the published algorithm with explicitly rounded float operations (no FMA contraction) makes the device
// projection agree with a reference built against that libm bit for bit.  tests/test_abi_cpu.py compiles this header for the host and
// compares it with the C library over random and special arguments.
#pragma once
#include <cstdint>
#include <cstring>

#if defined(__CUDA_ARCH__)
#define FD_HD __host__ __device__ __forceinline__
#define FD_ADD(a, b) __fadd_rn(a, b)
#define FD_SUB(a, b) __fsub_rn(a, b)
#define FD_MUL(a, b) __fmul_rn(a, b)
#define FD_DIV(a, b) __fdiv_rn(a, b)
#elif defined(__CUDACC__)
#define FD_HD __host__ __device__ inline
#define FD_ADD(a, b) ((a) + (b))
#define FD_SUB(a, b) ((a) - (b))
#define FD_MUL(a, b) ((a) * (b))
#define FD_DIV(a, b) ((a) / (b))
#else  // plain host compile (the CPU self-test); build with -ffp-contract=off
#define FD_HD inline
#define FD_ADD(a, b) ((a) + (b))
#define FD_SUB(a, b) ((a) - (b))
#define FD_MUL(a, b) ((a) * (b))
#define FD_DIV(a, b) ((a) / (b))
#endif

namespace fd {

FD_HD int32_t word(float x) {
#if defined(__CUDA_ARCH__)
  return __float_as_int(x);
#else
  int32_t i;
  memcpy(&i, &x, 4);
  return i;
#endif
}
FD_HD float from_word(int32_t i) {
#if defined(__CUDA_ARCH__)
  return __int_as_float(i);
#else
  float x;
  memcpy(&x, &i, 4);
  return x;
#endif
}

FD_HD float atanf(float x) {
  const float hi[4] = {4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f, 1.5707962513e+00f};
  const float lo[4] = {5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f, 7.5497894159e-08f};
  const float a0 = 3.3333334327e-01f, a1 = -2.0000000298e-01f, a2 = 1.4285714924e-01f, a3 = -1.1111110449e-01f, a4 = 9.0908870101e-02f,
              a5 = -7.6918758452e-02f, a6 = 6.6610731184e-02f, a7 = -5.8335702866e-02f, a8 = 4.9768779427e-02f, a9 = -3.6531571299e-02f,
              a10 = 1.6285819933e-02f;
  const int32_t hx = word(x), ix = hx & 0x7fffffff;
  int id;
  if (ix >= 0x4c000000) {  // |x| >= 2^25
    if (ix > 0x7f800000) return FD_ADD(x, x);
    const float r = FD_ADD(hi[3], lo[3]);
    return hx > 0 ? r : -r;
  }
  if (ix < 0x3ee00000) {  // |x| < 7/16
    if (ix < 0x31000000) return x;
    id = -1;
  } else {
    x = from_word(ix);
    if (ix < 0x3f980000) {
      if (ix < 0x3f300000) id = 0, x = FD_DIV(FD_SUB(FD_MUL(2.0f, x), 1.0f), FD_ADD(2.0f, x));
      else id = 1, x = FD_DIV(FD_SUB(x, 1.0f), FD_ADD(x, 1.0f));
    } else {
      if (ix < 0x401c0000) id = 2, x = FD_DIV(FD_SUB(x, 1.5f), FD_ADD(1.0f, FD_MUL(1.5f, x)));
      else id = 3, x = FD_DIV(-1.0f, x);
    }
  }
  const float z = FD_MUL(x, x), w = FD_MUL(z, z);
  float s1 = FD_ADD(a8, FD_MUL(w, a10));
  s1 = FD_ADD(a6, FD_MUL(w, s1)), s1 = FD_ADD(a4, FD_MUL(w, s1)), s1 = FD_ADD(a2, FD_MUL(w, s1)), s1 = FD_ADD(a0, FD_MUL(w, s1));
  s1 = FD_MUL(z, s1);
  float s2 = FD_ADD(a7, FD_MUL(w, a9));
  s2 = FD_ADD(a5, FD_MUL(w, s2)), s2 = FD_ADD(a3, FD_MUL(w, s2)), s2 = FD_ADD(a1, FD_MUL(w, s2));
  s2 = FD_MUL(w, s2);
  const float t = FD_MUL(x, FD_ADD(s1, s2));
  if (id < 0) return FD_SUB(x, t);
  const float r = FD_SUB(hi[id], FD_SUB(FD_SUB(t, lo[id]), x));
  return hx < 0 ? -r : r;
}

FD_HD float atan2f(float y, float x) {
  const float pi = 3.1415927410e+00f, pi_o_2 = 1.5707963705e+00f, pi_o_4 = 7.8539818525e-01f, pi_lo = -8.7422776573e-08f, tiny = 1.0e-30f;
  const int32_t hx = word(x), hy = word(y), ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
  if (ix > 0x7f800000 || iy > 0x7f800000) return FD_ADD(x, y);
  if (hx == 0x3f800000) return fd::atanf(y);
  const int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
  if (iy == 0) return m < 2 ? y : (m == 2 ? FD_ADD(pi, tiny) : FD_SUB(-pi, tiny));
  if (ix == 0) return hy < 0 ? FD_SUB(-pi_o_2, tiny) : FD_ADD(pi_o_2, tiny);
  if (ix == 0x7f800000) {
    if (iy == 0x7f800000) {
      const float q = FD_ADD(pi_o_4, tiny), q3 = FD_ADD(FD_MUL(3.0f, pi_o_4), tiny);
      return m == 0 ? q : (m == 1 ? -q : (m == 2 ? q3 : -q3));
    }
    return m == 0 ? 0.0f : (m == 1 ? -0.0f : (m == 2 ? FD_ADD(pi, tiny) : FD_SUB(-pi, tiny)));
  }
  if (iy == 0x7f800000) return hy < 0 ? FD_SUB(-pi_o_2, tiny) : FD_ADD(pi_o_2, tiny);
  const int k = (iy - ix) >> 23;
  float z;
  if (k > 60) z = FD_ADD(pi_o_2, FD_MUL(0.5f, pi_lo));
  else if (hx < 0 && k < -60) z = 0.0f;
  else z = fd::atanf(from_word(word(FD_DIV(y, x)) & 0x7fffffff));
  switch (m) {
    case 0: return z;
    case 1: return from_word(word(z) ^ (int32_t)0x80000000);
    case 2: return FD_SUB(pi, FD_SUB(z, pi_lo));
    default: return FD_SUB(FD_SUB(z, pi_lo), pi);
  }
}

}  // namespace fd
