/* Deterministic expf / logf shared by the CUDA sampler kernel and its C oracle.
 * Every operation is a single correctly-rounded IEEE-754 binary32 op (mul, add, fma, rint,
 * exact power-of-two scaling), so host and device produce bit-identical results and the sampled
 * permutations can be compared bit-exactly.  Polynomials: Cephes expf / logf.
 * Host translation units including this header must be built with -ffp-contract=off. */
#ifndef LR2_DET_MATH_H
#define LR2_DET_MATH_H
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define LR2_DM_FN __device__ __forceinline__
#define LR2_MUL(a, b) __fmul_rn((a), (b))
#define LR2_ADD(a, b) __fadd_rn((a), (b))
#define LR2_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#define LR2_BITS2F(i) __int_as_float(i)
#define LR2_F2BITS(f) __float_as_int(f)
#else
#define LR2_DM_FN static inline
#define LR2_MUL(a, b) ((a) * (b))
#define LR2_ADD(a, b) ((a) + (b))
#define LR2_FMA(a, b, c) fmaf((a), (b), (c))
static inline float lr2_bits2f_(int32_t i) { float f; memcpy(&f, &i, 4); return f; }
static inline int32_t lr2_f2bits_(float f) { int32_t i; memcpy(&i, &f, 4); return i; }
#define LR2_BITS2F(i) lr2_bits2f_(i)
#define LR2_F2BITS(f) lr2_f2bits_(f)
#endif

LR2_DM_FN float lr2_det_expf(float x) {
  if (x < -87.0f) return 0.0f;
  if (x > 88.0f) x = 88.0f;
  const float n = rintf(LR2_MUL(x, 1.44269504088896341f));
  float r = LR2_FMA(n, -0.693359375f, x);
  r = LR2_FMA(n, 2.12194440e-4f, r);
  float p = 1.9875691500e-4f;
  p = LR2_FMA(p, r, 1.3981999507e-3f);
  p = LR2_FMA(p, r, 8.3334519073e-3f);
  p = LR2_FMA(p, r, 4.1665795894e-2f);
  p = LR2_FMA(p, r, 1.6666665459e-1f);
  p = LR2_FMA(p, r, 5.0000001201e-1f);
  p = LR2_FMA(p, LR2_MUL(r, r), r);
  p = LR2_ADD(p, 1.0f);
  const int32_t e = (int32_t)n;  /* in [-126, 127] */
  return LR2_MUL(p, LR2_BITS2F((e + 127) << 23));
}

LR2_DM_FN float lr2_det_logf(float x) {
  if (!(x > 0.0f)) return -INFINITY;
  int32_t bits = LR2_F2BITS(x);
  int32_t e = ((bits >> 23) & 0xFF) - 127;
  float m = LR2_BITS2F((bits & 0x007FFFFF) | 0x3F800000); /* [1,2) */
  if (m > 1.41421356237f) { m = LR2_MUL(m, 0.5f); e += 1; }
  const float f = LR2_ADD(m, -1.0f);
  const float z = LR2_MUL(f, f);
  float y = 7.0376836292e-2f;
  y = LR2_FMA(y, f, -1.1514610310e-1f);
  y = LR2_FMA(y, f, 1.1676998740e-1f);
  y = LR2_FMA(y, f, -1.2420140846e-1f);
  y = LR2_FMA(y, f, 1.4249322787e-1f);
  y = LR2_FMA(y, f, -1.6668057665e-1f);
  y = LR2_FMA(y, f, 2.0000714765e-1f);
  y = LR2_FMA(y, f, -2.4999993993e-1f);
  y = LR2_FMA(y, f, 3.3333331174e-1f);
  y = LR2_MUL(LR2_MUL(y, f), z);
  const float fe = (float)e;
  y = LR2_FMA(fe, -2.12194440e-4f, y);
  y = LR2_FMA(z, -0.5f, y);
  float r = LR2_ADD(f, y);
  r = LR2_FMA(fe, 0.693359375f, r);
  return r;
}

#endif /* LR2_DET_MATH_H */
