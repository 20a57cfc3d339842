// Shared device helpers for the lr2ppo_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <math.h>

#include "../../include/lr2ppo_b200.h"

// Launch-error check used by every C-ABI entry point: never throws, never exits.
#define LR2_RETURN_LAUNCH()                                   \
  do {                                                        \
    cudaError_t e__ = cudaGetLastError();                     \
    return e__ == cudaSuccess ? LR2_OK : LR2_ERR_CUDA;        \
  } while (0)

// Every kernel launch is counted (bench.py reports the number inside its timed region).
extern "C" void lr2_note_launches(int n);
#define LR2_LAUNCHED(n) lr2_note_launches(n)

namespace lr2 {

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// exact (erf) GELU, matching torch.nn.GELU() / tencentpretrain act_fun.gelu
__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// Fast GELU for bf16 epilogues: erf by Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7, three orders of
// magnitude below bf16 resolution) with one MUFU.RCP and one MUFU.EX2; cdf and pdf share the exponential.
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float gelu_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = rcp_approx(fmaf(0.3275911f, z, 1.0f));
  const float poly = t * (0.254829592f + t * (-0.284496736f + t * (1.421413741f + t * (-1.453152027f + t * 1.061405429f))));
  const float e = ex2_approx(-z * z * 1.4426950408889634f);
  const float erf_abs = 1.0f - poly * e;            // erf(|x|/sqrt2)
  const float cdf = 0.5f * (1.0f + copysignf(erf_abs, x));
  return x * cdf;
}
__device__ __forceinline__ float gelu_fast_grad(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = rcp_approx(fmaf(0.3275911f, z, 1.0f));
  const float poly = t * (0.254829592f + t * (-0.284496736f + t * (1.421413741f + t * (-1.453152027f + t * 1.061405429f))));
  const float e = ex2_approx(-z * z * 1.4426950408889634f);   // = exp(-x^2/2)
  const float cdf = 0.5f * (1.0f + copysignf(1.0f - poly * e, x));
  return cdf + x * 0.39894228040143267794f * e;
}

// QuickGELU x * sigmoid(1.702 x) (CLIP-style blocks, ref: finetune/video_transformer.py:91-93)
__device__ __forceinline__ float qgelu(float x) { return x * rcp_approx(1.0f + ex2_approx(-1.702f * 1.4426950408889634f * x)); }
__device__ __forceinline__ float qgelu_grad(float x) {
  const float s = rcp_approx(1.0f + ex2_approx(-1.702f * 1.4426950408889634f * x));
  return s + 1.702f * x * s * (1.0f - s);
}
__device__ __forceinline__ float act_fwd(int act, float x) { return act == 1 ? qgelu(x) : gelu_fast(x); }
__device__ __forceinline__ float act_grad(int act, float x) { return act == 1 ? qgelu_grad(x) : gelu_fast_grad(x); }

// ---- Philox4x32-10 counter RNG (dropout masks are a pure function of
// (seed, site, element index), so backward regenerates them) -------------
struct Philox4 {
  uint32_t x, y, z, w;
};
__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  Philox4 r; r.x = c0; r.y = c1; r.z = c2; r.w = c3;
  return r;
}

// keep-mask for 4 consecutive elements starting at linear index idx4*4.
// keep iff u32 >= thresh where thresh = p * 2^32.
__device__ __forceinline__ uint32_t dropout_keep4(uint64_t seed, uint32_t site, uint64_t idx4, uint32_t thresh) {
  Philox4 r = philox4x32_10((uint32_t)idx4, (uint32_t)(idx4 >> 32), site, 0u,
                            (uint32_t)seed, (uint32_t)(seed >> 32));
  return (r.x >= thresh ? 1u : 0u) | (r.y >= thresh ? 2u : 0u) | (r.z >= thresh ? 4u : 0u) |
         (r.w >= thresh ? 8u : 0u);
}

// ---- 8-wide dropout stream (GEMM epilogues, LayerNorm backward, lr2_dropout_bf16) -------------------------------
// ONE Philox4x32-7 call yields 8 x 16 random bits: element j of the aligned group idx8 is kept iff its 16-bit field
// >= thresh16 = round(p * 65536).  The drop probability is therefore quantised to thresh16 / 65536 (|error| < 8e-6)
// and the survivors are scaled by exactly 65536 / (65536 - thresh16), so E[mult] == 1.  7 rounds is the smallest
// Philox4x32 variant that passes BigCrush (Salmon et al., SC'11); the shorter chain halves the epilogue ALU cost.
__host__ __device__ __forceinline__ uint32_t dropout_thresh16(float p) {
  float t = p * 65536.0f + 0.5f;
  if (t < 0.f) t = 0.f;
  if (t > 65535.f) t = 65535.f;
  return (uint32_t)t;
}
__host__ __device__ __forceinline__ float dropout_scale16(float p) {
  return p > 0.f ? 65536.0f / (65536.0f - (float)dropout_thresh16(p)) : 1.0f;
}
// m[j] = keep_j ? scale : 0 for the 8 elements [idx8*8, idx8*8 + 8)
__device__ __forceinline__ void dropout_mult8(uint64_t seed, uint32_t site, uint64_t idx8, uint32_t thresh16,
                                              float scale, float (&m)[8]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  uint32_t c0 = (uint32_t)idx8, c1 = (uint32_t)(idx8 >> 32), c2 = site, c3 = 0x38u;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += W0; k1 += W1;
  }
  m[0] = (c0 & 0xFFFFu) >= thresh16 ? scale : 0.f;  m[1] = (c0 >> 16) >= thresh16 ? scale : 0.f;
  m[2] = (c1 & 0xFFFFu) >= thresh16 ? scale : 0.f;  m[3] = (c1 >> 16) >= thresh16 ? scale : 0.f;
  m[4] = (c2 & 0xFFFFu) >= thresh16 ? scale : 0.f;  m[5] = (c2 >> 16) >= thresh16 ? scale : 0.f;
  m[6] = (c3 & 0xFFFFu) >= thresh16 ? scale : 0.f;  m[7] = (c3 >> 16) >= thresh16 ? scale : 0.f;
}

__host__ __device__ __forceinline__ uint32_t dropout_thresh(float p) {
  double t = (double)p * 4294967296.0;
  if (t < 0) t = 0;
  if (t > 4294967295.0) t = 4294967295.0;
  return (uint32_t)t;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

// 8 bf16 (one 16-byte vector) -> 8 floats
__device__ __forceinline__ void unpack8(const uint4& u, float (&x)[8]) {
  x[0] = __uint_as_float(u.x << 16); x[1] = __uint_as_float(u.x & 0xFFFF0000u);
  x[2] = __uint_as_float(u.y << 16); x[3] = __uint_as_float(u.y & 0xFFFF0000u);
  x[4] = __uint_as_float(u.z << 16); x[5] = __uint_as_float(u.z & 0xFFFF0000u);
  x[6] = __uint_as_float(u.w << 16); x[7] = __uint_as_float(u.w & 0xFFFF0000u);
}

}  // namespace lr2
