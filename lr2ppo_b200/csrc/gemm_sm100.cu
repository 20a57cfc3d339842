// lr2_gemm_bf16: persistent, warp-specialised tcgen05/TMEM GEMM for sm_100a.
//
//   D[M,N] = A[M,K] * B[N,K]^T      bf16 operands, fp32 accumulation in TMEM
//
// Operands arrive through TMA (128-byte swizzle) and may each be K-major
// (row-major [rows,K]) or MN-major (row-major [K,rows]) so that forward,
// dgrad and wgrad of every Linear on the LR2PPO hot path
// (reference finetune/ppo.py:154-170 Mlp, finetune/xit.py:103-147 FFN/Q/K/V/O)
// run without a transpose pass.  The epilogue fuses bias, exact-erf GELU,
// dropout (Philox), residual add and GELU-backward, writes bf16 or fp32, and can
// write the tile transposed (used by the skinny out_layer.fc1 GEMM where the
// 500 M-parameter weight is the 128-row "A" operand and the <=256 items are "N").
// Split-K writes fp32 slabs that lr2_splitk_reduce folds with the same epilogue.
//
// Roles per CTA (192 threads): warp0 = TMA producer, warp1 = TMEM alloc + MMA
// issuer (one elected lane), warps2-5 = epilogue (TMEM -> registers -> global).
// Two TMEM accumulator buffers let the epilogue of tile i overlap the MMAs of
// tile i+1.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <mutex>
#include <unordered_map>
#include <string>
#include <cstring>
#include <cstdlib>
#include "common.cuh"
#include "tc05.cuh"
#include "../../include/lr2ppo_b200.h"

namespace lr2 {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int UMMA_K = 16;
#ifndef LR2_EPI_WARPS
#define LR2_EPI_WARPS 16
#endif
constexpr int EPI_WARPS = LR2_EPI_WARPS;          // a multiple of 4: EPI_WARPS / 4 column parts per TMEM lane quadrant
constexpr int GEMM_THREADS = 64 + EPI_WARPS * 32;   // warp0 TMA, warp1 MMA, warps2-17 epilogue
constexpr int STG_PITCH = 32;                      // floats per staged row: dense 128-byte rows, XOR-swizzled chunks
// Per-warp epilogue staging tile: 32 rows x 32 fp32, 16-byte chunk c of row r stored at chunk (c ^ (r & 7)).  The
// row-per-lane writes (one row per lane) and the 4-lanes-per-row reads are both bank-conflict free without padding,
// which frees 9 KB per CTA: the 256 x 256 pair kernel gets a fifth smem stage.
__device__ __forceinline__ int stg_chunk(int r, int c) { return r * STG_PITCH + ((c ^ (r & 7)) << 2); }

struct GemmParams {
  int M, N, K;
  int splits;          // split-K factor (>=1)
  int kb_per_split;    // k-blocks per split
  int epi;             // LR2_EPI_*
  int transposed_out;  // output element (m,n) is written at C[n*ldc + m]
  int c_f32;           // output dtype: 1 = fp32, 0 = bf16
  void* C;
  long long ldc;
  bf16* C2;            // optional pre-activation output (EPI_BIAS_GELU), same shape/ld as C
  const float* bias;   // indexed by output column
  const bf16* aux;     // residual / pre-activation, same shape as output
  long long ldaux;
  float beta;          // fp32 out only: out += beta * C_old
  float drop_p;        // 0 = no dropout
  unsigned int drop_thresh;  // round(drop_p * 2^16) (host-computed, see dropout_mult8)
  float drop_scale;    // 65536 / (65536 - drop_thresh)
  unsigned long long seed;
  const unsigned long long* seed_dev;  // optional device-resident seed offset (CUDA-graph replay safe)
  unsigned int site;
  float* ws;           // split-K workspace [splits][out_rows*ldc]
  long long ws_slab;   // elements per slab
  int raster;          // 0: tiles strided over CTAs, m fastest; 1: contiguous chunk per CTA, n fastest
  int act;             // activation of the GELU epilogues: 0 = exact erf GELU, 1 = QuickGELU x*sigmoid(1.702x)
  int mode;            // EpiMode resolved on the host from (epi, act, drop_p, bias)
  int dbg;             // LR2_GEMM_DBG (profiling experiments only): 1 = skip the accumulator drain, 2 = skip the MMAs
  int direct;          // pair-kernel instantiation for plain bf16 outputs: 0 staged, 1 = 256-bit register stores, 2 = TMA store
  // LR2_EPI_ADAMW (fused wgrad + AdamW): C = fp32 parameter (in/out)
  float* adam_m; float* adam_v; bf16* adam_shadow; const float* adam_hyper; float adam_wd;
};

// kind::f16 instruction descriptor: D=f32 (1@4), A=B=bf16 (1@7, 1@10), majors @15/@16, N>>3 @17, M>>4 @24.
__host__ __device__ constexpr uint32_t make_idesc(int n, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

// -------------------------------------------------------------- epilogue --
// Host-resolved epilogue mode (GemmParams::mode): one warp-uniform switch instead of per-element conditionals.
enum EpiMode {
  EM_GENERIC = 0, EM_NONE, EM_BETA, EM_BIAS, EM_BIAS_GELU, EM_BIAS_GELU_DROP, EM_BIAS_RES, EM_BIAS_DROP_RES,
  EM_DGELU, EM_DGELU_DROP, EM_ADD
};

static int resolve_mode(int epi, int act, float drop_p, const float* bias, int c_f32, float beta) {
  switch (epi) {
    case LR2_EPI_NONE: return (c_f32 && beta != 0.f) ? EM_BETA : EM_NONE;
    case LR2_EPI_BIAS: return bias ? EM_BIAS : EM_GENERIC;
    case LR2_EPI_BIAS_GELU: return (bias && act == 0) ? (drop_p > 0.f ? EM_BIAS_GELU_DROP : EM_BIAS_GELU) : EM_GENERIC;
    case LR2_EPI_BIAS_DROP_RES: return bias ? (drop_p > 0.f ? EM_BIAS_DROP_RES : EM_BIAS_RES) : EM_GENERIC;
    case LR2_EPI_DGELU: return act == 0 ? (drop_p > 0.f ? EM_DGELU_DROP : EM_DGELU) : EM_GENERIC;
    case LR2_EPI_ADD: return EM_ADD;
    default: return EM_GENERIC;
  }
}

// One output row r, 8 consecutive output columns c..c+7 (c % 8 == 0, c + 8 <= ncols).
// epi_math8 loads bias / aux / old C and transforms v in registers (pre = bf16 pre-activation for C2);
// epi_write8 issues the stores.  Split so that the caller can interleave two independent groups.
// Output selection of one tile: the caller's epilogue, or (split-K) a raw fp32 partial into this split's slab.
// Kept separate from GemmParams so that the kernel parameter block is never copied per tile.
struct OutSel {
  void* C; int c_f32; int mode; int epi; float beta;
};

__device__ __forceinline__ void add_bias8(const GemmParams& p, float (&v)[8], int c) {
  const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + c));
  const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + c + 4));
  v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
  v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
}

template <bool ADAMW>
__device__ __forceinline__ void epi_math8(const GemmParams& p, const OutSel& o, float (&v)[8], uint4& pre, long long r, int c) {
  const long long off = r * p.ldc + c;
  if (!ADAMW && o.epi == LR2_EPI_NONE) {
    if (o.c_f32 && o.beta != 0.f) {
      const float4* cp = reinterpret_cast<const float4*>((const float*)o.C + off);
      float4 a = cp[0], b = cp[1];
      v[0] += o.beta * a.x; v[1] += o.beta * a.y; v[2] += o.beta * a.z; v[3] += o.beta * a.w;
      v[4] += o.beta * b.x; v[5] += o.beta * b.y; v[6] += o.beta * b.z; v[7] += o.beta * b.w;
    }
    return;
  }
  if constexpr (ADAMW) {
    // ref: tencentpretrain/utils/optimizers.py:374-402; acc = this tile of the weight gradient
    const float lr = p.adam_hyper[0], b1 = p.adam_hyper[1], b2 = p.adam_hyper[2], eps = p.adam_hyper[3],
                omb1 = p.adam_hyper[4], omb2 = p.adam_hyper[5], gs = p.adam_hyper[6], lrd = p.adam_hyper[7];
    float* P = (float*)o.C + off; float* Mm = p.adam_m + off; float* Vv = p.adam_v + off;
    const float4 p0 = *reinterpret_cast<const float4*>(P), p1 = *reinterpret_cast<const float4*>(P + 4);
    const float4 m0 = *reinterpret_cast<const float4*>(Mm), m1 = *reinterpret_cast<const float4*>(Mm + 4);
    const float4 v0 = *reinterpret_cast<const float4*>(Vv), v1 = *reinterpret_cast<const float4*>(Vv + 4);
    float pp[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
    float mm[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
    float vv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float g = v[i] * gs;
      mm[i] = mm[i] * b1 + g * omb1;
      vv[i] = vv[i] * b2 + (g * g) * omb2;
      const float denom = sqrtf(vv[i]) + eps;
      float pn = pp[i] - lr * (mm[i] / denom);
      if (p.adam_wd > 0.f) pn = pn - (lrd * p.adam_wd) * pn;
      v[i] = pn;
    }
    *reinterpret_cast<float4*>(Mm) = make_float4(mm[0], mm[1], mm[2], mm[3]);
    *reinterpret_cast<float4*>(Mm + 4) = make_float4(mm[4], mm[5], mm[6], mm[7]);
    *reinterpret_cast<float4*>(Vv) = make_float4(vv[0], vv[1], vv[2], vv[3]);
    *reinterpret_cast<float4*>(Vv + 4) = make_float4(vv[4], vv[5], vv[6], vv[7]);
    if (p.adam_shadow != nullptr) {
      uint4 u;
      u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
      u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
      *reinterpret_cast<uint4*>(p.adam_shadow + off) = u;
    }
    return;  // epi_write8 stores v (the new parameter) as fp32 into C
  } else {
  // Warp-uniform switch on the host-resolved mode; every hot case is straight-line code over the 8 elements.
  switch (o.mode) {
    case EM_BIAS: {
      add_bias8(p, v, c);
      return;
    }
    case EM_BIAS_GELU:
    case EM_BIAS_GELU_DROP: {
      add_bias8(p, v, c);
      // GELU is evaluated on the bf16-rounded pre-activation so that backward (which only has the stored bf16
      // copy) differentiates the same function.
      pre.x = pack_bf16x2(v[0], v[1]); pre.y = pack_bf16x2(v[2], v[3]);
      pre.z = pack_bf16x2(v[4], v[5]); pre.w = pack_bf16x2(v[6], v[7]);
      float x[8];
      unpack8(pre, x);
      if (o.mode == EM_BIAS_GELU) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = gelu_fast(x[i]);
      } else {
        float m[8];
        dropout_mult8(p.seed + (p.seed_dev ? *p.seed_dev : 0ull), p.site, (uint64_t)off >> 3, p.drop_thresh,
                      p.drop_scale, m);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = gelu_fast(x[i]) * m[i];
      }
      return;
    }
    case EM_BIAS_RES:
    case EM_BIAS_DROP_RES: {
      add_bias8(p, v, c);
      float a[8];
      unpack8(*reinterpret_cast<const uint4*>(p.aux + r * p.ldaux + c), a);
      if (o.mode == EM_BIAS_RES) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] += a[i];
      } else {
        float m[8];
        dropout_mult8(p.seed + (p.seed_dev ? *p.seed_dev : 0ull), p.site, (uint64_t)off >> 3, p.drop_thresh,
                      p.drop_scale, m);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = fmaf(v[i], m[i], a[i]);
      }
      return;
    }
    case EM_DGELU:
    case EM_DGELU_DROP: {
      float a[8];
      unpack8(*reinterpret_cast<const uint4*>(p.aux + r * p.ldaux + c), a);
      if (o.mode == EM_DGELU) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] *= gelu_fast_grad(a[i]);
      } else {
        float m[8];
        dropout_mult8(p.seed + (p.seed_dev ? *p.seed_dev : 0ull), p.site, (uint64_t)off >> 3, p.drop_thresh,
                      p.drop_scale, m);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] *= gelu_fast_grad(a[i]) * m[i];
      }
      return;
    }
    case EM_ADD: {
      float a[8];
      unpack8(*reinterpret_cast<const uint4*>(p.aux + r * p.ldaux + c), a);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] += a[i];
      return;
    }
    default:
      break;
  }
  // Generic path (QuickGELU variants): same semantics, runtime-selected activation.
  if (p.bias != nullptr && (o.epi == LR2_EPI_BIAS || o.epi == LR2_EPI_BIAS_GELU || o.epi == LR2_EPI_BIAS_DROP_RES))
    add_bias8(p, v, c);
  float a[8];
  if (o.epi == LR2_EPI_BIAS_DROP_RES || o.epi == LR2_EPI_DGELU || o.epi == LR2_EPI_ADD)
    unpack8(*reinterpret_cast<const uint4*>(p.aux + r * p.ldaux + c), a);
  float m[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
  if (p.drop_p > 0.f)
    dropout_mult8(p.seed + (p.seed_dev ? *p.seed_dev : 0ull), p.site, (uint64_t)off >> 3, p.drop_thresh, p.drop_scale, m);
  if (o.epi == LR2_EPI_BIAS_GELU) {
    pre.x = pack_bf16x2(v[0], v[1]); pre.y = pack_bf16x2(v[2], v[3]);
    pre.z = pack_bf16x2(v[4], v[5]); pre.w = pack_bf16x2(v[6], v[7]);
    float x[8];
    unpack8(pre, x);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = act_fwd(p.act, x[i]) * m[i];
  } else if (o.epi == LR2_EPI_BIAS_DROP_RES) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = fmaf(v[i], m[i], a[i]);
  } else if (o.epi == LR2_EPI_DGELU) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] *= act_grad(p.act, a[i]) * m[i];
  } else if (o.epi == LR2_EPI_ADD) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] += a[i];
  }
  }
}
__device__ __forceinline__ void epi_write8(const GemmParams& p, const OutSel& o, const float (&v)[8], const uint4& pre,
                                           long long r, int c) {
  const long long off = r * p.ldc + c;
  if (o.epi == LR2_EPI_BIAS_GELU && p.C2 != nullptr) *reinterpret_cast<uint4*>(p.C2 + off) = pre;
  if (o.c_f32) {
    float4* cp = reinterpret_cast<float4*>((float*)o.C + off);
    cp[0] = make_float4(v[0], v[1], v[2], v[3]);
    cp[1] = make_float4(v[4], v[5], v[6], v[7]);
  } else {
    uint4 u;
    u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
    u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>((bf16*)o.C + off) = u;
  }
}
__device__ __forceinline__ void epi_store8(const GemmParams& p, const OutSel& o, float (&v)[8], long long r, int c) {
  uint4 pre = make_uint4(0, 0, 0, 0);
  epi_math8<false>(p, o, v, pre, r, c);
  epi_write8(p, o, v, pre, r, c);
}

// Scalar epilogue (transposed tiles and ragged edges). No dropout on this path.
__device__ __forceinline__ void epi_store1(const GemmParams& p, const OutSel& o, float v, long long r, int c) {
  const long long off = r * p.ldc + c;
  if (o.epi == LR2_EPI_NONE) {
    if (o.c_f32 && o.beta != 0.f) v += o.beta * ((const float*)o.C)[off];
  } else {
    if (p.bias != nullptr && (o.epi == LR2_EPI_BIAS || o.epi == LR2_EPI_BIAS_GELU || o.epi == LR2_EPI_BIAS_DROP_RES))
      v += __ldg(p.bias + c);
    float a = 0.f;
    if (o.epi == LR2_EPI_BIAS_DROP_RES || o.epi == LR2_EPI_DGELU || o.epi == LR2_EPI_ADD)
      a = __bfloat162float(p.aux[r * p.ldaux + c]);
    if (o.epi == LR2_EPI_BIAS_GELU) {
      if (p.C2 != nullptr) p.C2[off] = __float2bfloat16(v);
      v = act_fwd(p.act, __bfloat162float(__float2bfloat16(v)));
    } else if (o.epi == LR2_EPI_BIAS_DROP_RES || o.epi == LR2_EPI_ADD) {
      v += a;
    } else if (o.epi == LR2_EPI_DGELU) {
      v *= act_grad(p.act, a);
    }
  }
  if (o.c_f32) ((float*)o.C)[off] = v;
  else ((bf16*)o.C)[off] = __float2bfloat16(v);
}

// Fast path of the accumulator drain: one whole 32-row x 32-column chunk, staged in `stg` (row per lane), is pushed
// through the host-resolved epilogue MODE with every warp-uniform decision taken outside the row loop: bias (a
// per-column quantity) is loaded once per chunk, addresses advance by a constant stride, and the four row groups are
// fully unrolled so that their aux / C loads and Philox chains overlap.
template <int MODE, bool F32>
__device__ __forceinline__ void chunk_rows(const GemmParams& q, const OutSel& o, const float* __restrict__ stg, int lane,
                                           int m_base, int n_base) {
  constexpr int RPI = 8;                                   // rows per iteration (4 lanes x 8 columns per row)
  const int cg = (lane & 3) * 8;
  const int r0 = lane >> 2;
  const int n = n_base + cg;
  constexpr bool HAS_BIAS = MODE == EM_BIAS || MODE == EM_BIAS_GELU || MODE == EM_BIAS_GELU_DROP ||
                            MODE == EM_BIAS_RES || MODE == EM_BIAS_DROP_RES;
  constexpr bool HAS_AUX = MODE == EM_BIAS_RES || MODE == EM_BIAS_DROP_RES || MODE == EM_DGELU ||
                           MODE == EM_DGELU_DROP || MODE == EM_ADD;
  constexpr bool HAS_DROP = MODE == EM_BIAS_GELU_DROP || MODE == EM_BIAS_DROP_RES || MODE == EM_DGELU_DROP;
  float bias[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if constexpr (HAS_BIAS) {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(q.bias + n));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(q.bias + n + 4));
    bias[0] = b0.x; bias[1] = b0.y; bias[2] = b0.z; bias[3] = b0.w;
    bias[4] = b1.x; bias[5] = b1.y; bias[6] = b1.z; bias[7] = b1.w;
  }
  unsigned long long seed = 0;
  if constexpr (HAS_DROP) seed = q.seed + (q.seed_dev ? *q.seed_dev : 0ull);
  const long long off0 = (long long)(m_base + r0) * q.ldc + n;
  const long long step = (long long)RPI * q.ldc;
  const bf16* aux = nullptr;
  long long astep = 0;
  if constexpr (HAS_AUX) { aux = q.aux + (long long)(m_base + r0) * q.ldaux + n; astep = (long long)RPI * q.ldaux; }
  const bool want_pre = (MODE == EM_BIAS_GELU || MODE == EM_BIAS_GELU_DROP) && q.C2 != nullptr;
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const long long off = off0 + it * step;
    const int rs = it * RPI + r0;
    const float4 x0 = *reinterpret_cast<const float4*>(stg + stg_chunk(rs, cg >> 2));
    const float4 x1 = *reinterpret_cast<const float4*>(stg + stg_chunk(rs, (cg >> 2) + 1));
    float v[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
    float a[8];
    if constexpr (HAS_AUX) unpack8(*reinterpret_cast<const uint4*>(aux + it * astep), a);
    float m[8];
    if constexpr (HAS_DROP) dropout_mult8(seed, q.site, (uint64_t)off >> 3, q.drop_thresh, q.drop_scale, m);
    if constexpr (HAS_BIAS) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] += bias[i];
    }
    if constexpr (MODE == EM_BETA) {
      const float4* cp = reinterpret_cast<const float4*>((const float*)o.C + off);
      const float4 c0 = cp[0], c1 = cp[1];
      v[0] = fmaf(o.beta, c0.x, v[0]); v[1] = fmaf(o.beta, c0.y, v[1]); v[2] = fmaf(o.beta, c0.z, v[2]);
      v[3] = fmaf(o.beta, c0.w, v[3]); v[4] = fmaf(o.beta, c1.x, v[4]); v[5] = fmaf(o.beta, c1.y, v[5]);
      v[6] = fmaf(o.beta, c1.z, v[6]); v[7] = fmaf(o.beta, c1.w, v[7]);
    } else if constexpr (MODE == EM_BIAS_GELU || MODE == EM_BIAS_GELU_DROP) {
      // GELU of the bf16-rounded pre-activation: backward only has the stored bf16 copy
      uint4 pre;
      pre.x = pack_bf16x2(v[0], v[1]); pre.y = pack_bf16x2(v[2], v[3]);
      pre.z = pack_bf16x2(v[4], v[5]); pre.w = pack_bf16x2(v[6], v[7]);
      if (want_pre) *reinterpret_cast<uint4*>(q.C2 + off) = pre;
      float x[8];
      unpack8(pre, x);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = HAS_DROP ? gelu_fast(x[i]) * m[i] : gelu_fast(x[i]);
    } else if constexpr (MODE == EM_BIAS_RES || MODE == EM_ADD) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] += a[i];
    } else if constexpr (MODE == EM_BIAS_DROP_RES) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = fmaf(v[i], m[i], a[i]);
    } else if constexpr (MODE == EM_DGELU || MODE == EM_DGELU_DROP) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] *= HAS_DROP ? gelu_fast_grad(a[i]) * m[i] : gelu_fast_grad(a[i]);
    }
    if constexpr (F32) {
      float4* cp = reinterpret_cast<float4*>((float*)o.C + off);
      cp[0] = make_float4(v[0], v[1], v[2], v[3]);
      cp[1] = make_float4(v[4], v[5], v[6], v[7]);
    } else {
      uint4 u;
      u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
      u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
      *reinterpret_cast<uint4*>((bf16*)o.C + off) = u;
    }
  }
}

// ---- direct drain (round 2): TMEM -> registers -> global, no shared-memory staging --------------------------------
// tcgen05.ld.32x32b hands every lane ONE output row (32 consecutive fp32 columns).  The staged path above re-shapes
// that through shared memory into 4-lanes-per-row groups so that a warp store covers whole 64-byte row pieces; with
// the 256-bit global stores of sm_100 (STG.E.ENL2.256) a lane moves 16 bf16 = one full 32-byte sector of its row per
// instruction, so sector efficiency is the same and the staging round trip -- 8 STS.128 + 8 LDS.128 per lane and
// chunk -- disappears.  Compiled ONLY into the DIRECT instantiation of the pair kernel, which the host selects for
// plain (EM_NONE) untransposed bf16 outputs: measured on B200 it wins there (9408x3072x768: 47.1 vs 50.6 us) and
// loses wherever the epilogue carries math (GELU +14 %, dropout +10 %; profiles/r02_gemm_shapes_direct_all.txt).
// Keeping it out of every other instantiation matters: inlined next to the staged modes it pushed all GEMM kernels
// over their 96-register budget (124-232 bytes of spills, ~10 % on every shape).
__device__ __forceinline__ void st_global_256(void* ptr, const uint32_t (&u)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(ptr), "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]),
               "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7])
               : "memory");
}

__device__ __forceinline__ void chunk_direct_plain(const GemmParams& q, uint32_t taddr_c, int lane, int m_base, int n_base) {
  uint32_t r[32];
  tmem_ld32(taddr_c, r);
  bf16* dst = (bf16*)q.C + (long long)(m_base + lane) * q.ldc + n_base;
  tmem_ld_wait();
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    uint32_t u[8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
      u[i] = pack_bf16x2(__uint_as_float(r[16 * h + 2 * i]), __uint_as_float(r[16 * h + 2 * i + 1]));
    st_global_256(dst + 16 * h, u);
  }
}

// ---- TMA-store drain (DIRECT == 2) ------------------------------------------------------------------------------
// The warp's 32 rows x 64 bf16 columns of the tile are packed into its 4 KB staging buffer in the 128-byte-swizzle
// layout a {64, 32} box of the output tensor map expects (16-byte chunk c of row r at chunk c ^ (r & 7): the same
// conflict-free pattern stg_chunk uses), then ONE lane hands the box to the TMA unit (cp.async.bulk.tensor ...
// global.shared::cta, SASS UTMASTG).  The epilogue warps never wait on their own global stores, the TMA unit writes
// whole 128-byte lines, and boxes that stick out of the tensor are clipped by the hardware, so this path has no
// ragged-edge branch at all.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 32 accumulator columns [c32*32, c32*32+32) of this lane's row -> bf16 -> chunks c32*4 .. c32*4+3 of the row
__device__ __forceinline__ void stage_bf16_swizzled(uint8_t* stg, uint32_t taddr_c, int lane, int c32) {
  uint32_t r[32];
  tmem_ld32(taddr_c, r);
  tmem_ld_wait();
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    uint4 u;
    u.x = pack_bf16x2(__uint_as_float(r[8 * g + 0]), __uint_as_float(r[8 * g + 1]));
    u.y = pack_bf16x2(__uint_as_float(r[8 * g + 2]), __uint_as_float(r[8 * g + 3]));
    u.z = pack_bf16x2(__uint_as_float(r[8 * g + 4]), __uint_as_float(r[8 * g + 5]));
    u.w = pack_bf16x2(__uint_as_float(r[8 * g + 6]), __uint_as_float(r[8 * g + 7]));
    *reinterpret_cast<uint4*>(stg + lane * 128 + (((c32 * 4 + g) ^ (lane & 7)) << 4)) = u;
  }
}

// One warp's share (32 rows x 64 columns starting at column c0 of the tile) of a 256-wide accumulator tile.
__device__ __forceinline__ void drain_warp_tma(const CUtensorMap* tmap_c, uint8_t* stg, uint32_t taddr, int c0, int lane,
                                               int m_base, int n_base) {
  if (lane == 0) tma_store_wait_read();      // the previous box of this warp has left the staging buffer
  __syncwarp();
  stage_bf16_swizzled(stg, taddr + (uint32_t)c0, lane, 0);
  stage_bf16_swizzled(stg, taddr + (uint32_t)(c0 + 32), lane, 1);
  fence_proxy_async_smem();                  // generic-proxy writes -> visible to the async proxy (TMA)
  __syncwarp();
  if (lane == 0) {
    tma_store_2d(tmap_c, stg, n_base, m_base);   // rows >= M / columns >= N of the box are clipped by the TMA unit
    tma_store_commit();
  }
}

// Returns false when (mode, output type) has no fast instantiation; the caller then takes the checked generic loop.
__device__ __forceinline__ bool chunk_fast(const GemmParams& q, const OutSel& o, const float* stg, int lane, int m_base,
                                           int n_base) {
  if (o.c_f32) {
    if (o.mode == EM_NONE) { chunk_rows<EM_NONE, true>(q, o, stg, lane, m_base, n_base); return true; }
    if (o.mode == EM_BETA) { chunk_rows<EM_BETA, true>(q, o, stg, lane, m_base, n_base); return true; }
    return false;
  }
  switch (o.mode) {
    case EM_NONE: chunk_rows<EM_NONE, false>(q, o, stg, lane, m_base, n_base); return true;
    case EM_BIAS: chunk_rows<EM_BIAS, false>(q, o, stg, lane, m_base, n_base); return true;
    case EM_BIAS_GELU: chunk_rows<EM_BIAS_GELU, false>(q, o, stg, lane, m_base, n_base); return true;
    case EM_BIAS_GELU_DROP: chunk_rows<EM_BIAS_GELU_DROP, false>(q, o, stg, lane, m_base, n_base); return true;
    case EM_BIAS_RES: chunk_rows<EM_BIAS_RES, false>(q, o, stg, lane, m_base, n_base); return true;
    case EM_BIAS_DROP_RES: chunk_rows<EM_BIAS_DROP_RES, false>(q, o, stg, lane, m_base, n_base); return true;
    case EM_DGELU: chunk_rows<EM_DGELU, false>(q, o, stg, lane, m_base, n_base); return true;
    case EM_DGELU_DROP: chunk_rows<EM_DGELU_DROP, false>(q, o, stg, lane, m_base, n_base); return true;
    case EM_ADD: chunk_rows<EM_ADD, false>(q, o, stg, lane, m_base, n_base); return true;
    default: return false;
  }
}

// ---------------------------------------------------------------- kernel --
template <int BN>
struct SmemLayout {
  static constexpr int A_BYTES = BM * BK * 2;   // 16 KB
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 3 : (BN == 128 ? 5 : 6);
  static constexpr int BAR_BYTES = 256;
  static constexpr int STG_BYTES = EPI_WARPS * 32 * STG_PITCH * 4;      // epilogue staging (fp32), per warp 32 rows
  static constexpr int TOTAL = STAGES * STAGE_BYTES + BAR_BYTES + STG_BYTES + 1024;  // +1024 for manual alignment
};

// Checked generic path of the drain (ragged tile edges, fp32 outputs with fused math, QuickGELU, fused AdamW): one
// staged 32 x 32 chunk through epi_math8 / epi_store1 with per-group bounds checks.  Deliberately NOT inlined for the
// plain kernels: its many warp-uniform conditions otherwise get hoisted into the hot per-tile path (ncu showed ~150
// predicate-packing instructions per tile per warp).
template <bool ADAMW>
__device__ __forceinline__ void chunk_checked_body(const GemmParams& q, const OutSel& o, const float* stg, int lane,
                                                   int m_base, int n_base) {
  constexpr int TPR = 4, RPI = 8;
  const int cg = (lane % TPR) * 8;
#pragma unroll 1
  for (int rr = 0; rr < 32; rr += 2 * RPI) {
    const int rl0 = rr + lane / TPR, rl1 = rl0 + RPI;
    const int m0 = m_base + rl0, m1 = m_base + rl1;
    const int n = n_base + cg;
    const bool full = (n + 8 <= q.N);
    if (full && m1 < q.M) {
      // two independent 8-wide groups in flight (rows rl0 and rl1)
      const float4 x0 = *reinterpret_cast<const float4*>(stg + stg_chunk(rl0, cg >> 2));
      const float4 x1 = *reinterpret_cast<const float4*>(stg + stg_chunk(rl0, (cg >> 2) + 1));
      const float4 y0 = *reinterpret_cast<const float4*>(stg + stg_chunk(rl1, cg >> 2));
      const float4 y1 = *reinterpret_cast<const float4*>(stg + stg_chunk(rl1, (cg >> 2) + 1));
      float v0[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
      float v1[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
      uint4 p0 = make_uint4(0, 0, 0, 0), p1 = make_uint4(0, 0, 0, 0);
      if constexpr (ADAMW) {   // HBM-bound: one group at a time keeps the register budget
        epi_math8<true>(q, o, v0, p0, m0, n);
        epi_write8(q, o, v0, p0, m0, n);
        epi_math8<true>(q, o, v1, p1, m1, n);
        epi_write8(q, o, v1, p1, m1, n);
      } else {
        epi_math8<false>(q, o, v0, p0, m0, n);
        epi_math8<false>(q, o, v1, p1, m1, n);
        epi_write8(q, o, v0, p0, m0, n);
        epi_write8(q, o, v1, p1, m1, n);
      }
    } else {
#pragma unroll 1
      for (int h2 = 0; h2 < 2; ++h2) {
        const int row_l = h2 ? rl1 : rl0;
        const int m = m_base + row_l;
        if (m < q.M && n < q.N) {
          const float4 x0 = *reinterpret_cast<const float4*>(stg + stg_chunk(row_l, cg >> 2));
          const float4 x1 = *reinterpret_cast<const float4*>(stg + stg_chunk(row_l, (cg >> 2) + 1));
          float v[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
          if (full) {
            if constexpr (ADAMW) {
              uint4 pz = make_uint4(0, 0, 0, 0);
              epi_math8<true>(q, o, v, pz, m, n);
              epi_write8(q, o, v, pz, m, n);
            } else {
              epi_store8(q, o, v, m, n);
            }
          } else {
#pragma unroll 1
            for (int i = 0; i < 8; ++i)
              if (n + i < q.N) epi_store1(q, o, v[i], m, n + i);
          }
        }
      }
    }
  }
}
__device__ __noinline__ void chunk_checked(const GemmParams& q, const OutSel o, const float* stg, int lane, int m_base,
                                           int n_base) {
  chunk_checked_body<false>(q, o, stg, lane, m_base, n_base);
}
// transposed tile (skinny GEMMs): lane = output column m, registers = output rows; a warp store writes 32
// consecutive columns of one output row.
__device__ __noinline__ void chunk_transposed(const GemmParams& q, const OutSel o, uint32_t taddr_c, int lane, int m_base,
                                              int n_base) {
  const int m = m_base + lane;
  uint32_t r[32];
  tmem_ld32(taddr_c, r);
  tmem_ld_wait();
  if (m < q.M) {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (n_base + i < q.N) epi_store1(q, o, __uint_as_float(r[i]), n_base + i, m);
  }
}

// Drain one accumulator tile (this warp's TMEM lane quadrant and column part): TMEM -> registers -> per-warp smem
// staging -> coalesced 8-wide groups through the fused epilogue.  Shared by the 1-CTA and 2-CTA kernels.
template <int BN, bool ADAMW, int DIRECT = 0>
__device__ __forceinline__ void drain_tile(const GemmParams& q, const OutSel& o, uint32_t taddr, int m_base, int nt,
                                           float* stg, int lane, int half) {
  constexpr int PARTS = EPI_WARPS / 4;
  constexpr int COLS_PER_WARP = (BN / PARTS < 32) ? 32 : BN / PARTS;   // BN=64, 16 warps: only parts 0,1 have columns
  constexpr int CW = 32;                                        // chunk width
#pragma unroll 1
  for (int c0 = half * COLS_PER_WARP; c0 < (half + 1) * COLS_PER_WARP && c0 < BN; c0 += CW) {
    const int n_base = nt * BN + c0;
    if (n_base >= q.N) break;  // warp-uniform
    if (q.transposed_out) {
      chunk_transposed(q, o, taddr + (uint32_t)c0, lane, m_base, n_base);
      continue;
    }
    if constexpr (DIRECT == 1) {
      // host guarantees: plain epilogue, untransposed bf16 output, 32-byte aligned rows, no split-K
      if (m_base + 32 <= q.M && n_base + 32 <= q.N) {
        chunk_direct_plain(q, taddr + (uint32_t)c0, lane, m_base, n_base);
        continue;
      }
    }
    {
      uint32_t r[32];
      tmem_ld32(taddr + (uint32_t)c0, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 8; ++i)
        *reinterpret_cast<float4*>(stg + stg_chunk(lane, i)) =
            make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]),
                        __uint_as_float(r[4 * i + 3]));
    }
    __syncwarp();
    if constexpr (ADAMW) {
      chunk_checked_body<true>(q, o, stg, lane, m_base, n_base);
    } else if constexpr (DIRECT == 1) {
      chunk_checked(q, o, stg, lane, m_base, n_base);          // ragged edge chunks only
    } else {
      if (!(m_base + 32 <= q.M && n_base + 32 <= q.N && chunk_fast(q, o, stg, lane, m_base, n_base)))
        chunk_checked(q, o, stg, lane, m_base, n_base);
    }
    __syncwarp();
  }
}

template <int BN, bool A_MN, bool B_MN, bool ADAMW>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
            const __grid_constant__ GemmParams p) {
  using L = SmemLayout<BN>;
  constexpr int STAGES = L::STAGES;
  constexpr int NBUF = (BN <= 128) ? 4 : 2;          // TMEM accumulator buffers (epilogue of tile i overlaps MMAs of i+1..)
  constexpr uint32_t TMEM_COLS = (NBUF * BN <= 256) ? 256 : 512;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared space
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * L::STAGE_BYTES);
  uint64_t* full_bar = bars;                  // [STAGES]
  uint64_t* empty_bar = bars + STAGES;        // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;    // [NBUF]
  uint64_t* tempty_bar = bars + 2 * STAGES + NBUF;  // [NBUF]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 2 * NBUF);
  float* stg_all = reinterpret_cast<float*>(smem + STAGES * L::STAGE_BYTES + L::BAR_BYTES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_tiles = (p.M + BM - 1) / BM;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int num_kb = (p.K + BK - 1) / BK;
  const int total_work = m_tiles * n_tiles * p.splits;
  // work range of this CTA (identical in all three roles)
  int w_begin, w_end, w_step;
  if (p.raster == 0) { w_begin = blockIdx.x; w_end = total_work; w_step = gridDim.x; }
  else {
    const int per = (total_work + gridDim.x - 1) / gridDim.x;
    w_begin = blockIdx.x * per; w_end = min(total_work, w_begin + per); w_step = 1;
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
      for (int b = 0; b < NBUF; ++b) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], EPI_WARPS); }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ======================= TMA producer =======================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = w_begin; w < w_end; w += w_step) {
        const int split = w % p.splits;
        const int t = w / p.splits;
        const int mt = p.raster ? t / n_tiles : t % m_tiles, nt = p.raster ? t % n_tiles : t / m_tiles;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(num_kb, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], L::STAGE_BYTES);
          uint8_t* sa = smem + stage * L::STAGE_BYTES;
          uint8_t* sb = sa + L::A_BYTES;
          if (A_MN) {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j)
              tma_load_2d(sa + j * (64 * BK * 2), &tmap_a, &full_bar[stage], mt * BM + j * 64, kb * BK);
          } else {
            tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * BK, mt * BM);
          }
          if (B_MN) {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_2d(sb + j * (64 * BK * 2), &tmap_b, &full_bar[stage], nt * BN + j * 64, kb * BK);
          } else {
            tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * BK, nt * BN);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer =========================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BN, A_MN, B_MN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int w = w_begin; w < w_end; w += w_step, ++it) {
        const int split = w % p.splits;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(num_kb, kb0 + p.kb_per_split);
        const int buf = it & (NBUF - 1);
        const uint32_t acc_phase = (uint32_t)(it / NBUF) & 1u;
        mbar_wait(&tempty_bar[buf], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(buf * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * L::STAGE_BYTES);
          const uint32_t sb = sa + L::A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // K-major: 16 bf16 = 32 B along the swizzled 128-B row.
            // MN-major: 16 k-rows = two 8-row (1024 B) swizzle atoms.
            const uint64_t ad = A_MN ? make_sdesc(sa + k * 2048, 64 * BK * 2, 1024) : make_sdesc(sa + k * 32, 16, 1024);
            const uint64_t bd = B_MN ? make_sdesc(sb + k * 2048, 64 * BK * 2, 1024) : make_sdesc(sb + k * 32, 16, 1024);
            umma_bf16(tmem_d, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[buf]);
      }
    }
  } else {
    // ======================= epilogue warps =====================
    // warp -> (TMEM lane quadrant, column half). Per 64-column chunk: TMEM -> registers -> per-warp smem
    // staging (row per lane), then re-read with 8 lanes per row so that bias / aux / C / C2 accesses are
    // 128-byte coalesced and every lane has 8 independent elements in flight for the GELU / Philox math.
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;                             // column part 0..3
    float* stg = stg_all + (size_t)(warp - 2) * 32 * STG_PITCH;
    int it = 0;
    for (int w = w_begin; w < w_end; w += w_step, ++it) {
      const int split = w % p.splits;
      const int t = w / p.splits;
      const int mt = p.raster ? t / n_tiles : t % m_tiles, nt = p.raster ? t % n_tiles : t / m_tiles;
      const int buf = it & (NBUF - 1);
      const uint32_t acc_phase = (uint32_t)(it / NBUF) & 1u;
      mbar_wait(&tfull_bar[buf], acc_phase);
      tc_fence_after();
      const int m_base = mt * BM + quad * 32;
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * BN);
      OutSel o{p.C, p.c_f32, p.mode, p.epi, p.beta};
      if (p.splits > 1) {  // raw fp32 partial into this split's slab
        o.C = p.ws + (long long)split * p.ws_slab; o.c_f32 = 1; o.mode = EM_NONE; o.epi = LR2_EPI_NONE; o.beta = 0.f;
      }
      drain_tile<BN, ADAMW>(p, o, taddr, m_base, nt, stg, lane, half);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[buf]);
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------ 2-CTA pair kernel --
// cta_group::2 variant for the large GEMMs: a cluster of two CTAs (one TPC) computes a 256 x BN tile with
// tcgen05.mma.cta_group::2 (UMMA_M = 256).  CTA r of the pair loads rows [r*128, r*128+128) of the A tile and rows
// [r*BN/2, (r+1)*BN/2) of the B tile into its own shared memory and owns accumulator rows r*128.. in its own TMEM, so
// each SM receives (128 + BN/2) x 64 operand elements per k-block for 128 x BN outputs: for BN = 256 that is half the
// L2->SMEM bytes per FLOP of the single-CTA 128x128 tile, which is what bounds that kernel.
// Protocol (one mbarrier set per CTA at identical smem offsets):
//   full[s]   leader only : count 1, leader's producer arrives with expect_tx for BOTH CTAs' bytes; the peer's TMA
//                           signals the leader's barrier (shared::cluster address of rank 0)
//   empty[s]  both        : count 1, released by the leader's MMA thread with a multicast tcgen05.commit
//   tfull[b]  both        : count 1, multicast tcgen05.commit after the last k-block of a tile
//   tempty[b] leader only : count 2*EPI_WARPS, the epilogue warps of both CTAs arrive (remote arrive from the peer)
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Execution-only rendezvous of the pair (kernel end: neither CTA may exit or free TMEM while its peer still reads its
// smem / TMEM).  No data is published through it, so the arrive is relaxed: a .release arrive would first wait for the
// L2 acknowledgement of every output store of the CTA.
__device__ __forceinline__ void cluster_sync_exec() {
  asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_rank(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
// Arrive on a barrier of the pair's leader CTA.  Default semantics (.release.cta), as in the single-CTA kernel: the
// only thing the MMA thread consumes after this arrive is the TMEM accumulator buffer, whose reads are ordered by
// tcgen05.wait::ld + tcgen05.fence::before_thread_sync.  A .release.cluster arrive made every epilogue warp wait for
// the L2 acknowledgement of all its output stores first (MEMBAR.ALL.GPU + ERRBAR: 17 % of the kernel's stall
// samples in profiles/r02_gemm_plain_sass.md) before the accumulator buffer was handed back.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, uint32_t leader_bar, int c0,
                                                 int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}

template <int BN>   // BN = N of the pair tile; each CTA stages BN/2 rows of B
struct SmemLayout2 {
  static constexpr int HB = BN / 2;
  static constexpr int A_BYTES = BM * BK * 2;   // 16 KB
  static constexpr int B_BYTES = HB * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_BYTES = 256;
  static constexpr int STG_BYTES = EPI_WARPS * 32 * STG_PITCH * 4;
  static constexpr int STAGES = (232448 - 1024 - BAR_BYTES - STG_BYTES) / STAGE_BYTES > 8
                                    ? 8 : (232448 - 1024 - BAR_BYTES - STG_BYTES) / STAGE_BYTES;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + BAR_BYTES + STG_BYTES + 1024;
};

template <int BN, bool A_MN, bool B_MN, int DIRECT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
             const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ GemmParams p) {
  using L = SmemLayout2<BN>;
  constexpr int STAGES = L::STAGES;
  constexpr int HB = L::HB;
  constexpr int NBUF = (BN <= 128) ? 4 : 2;
  constexpr uint32_t TMEM_COLS = (NBUF * BN <= 256) ? 256 : 512;
  static_assert(NBUF * BN <= 512, "accumulator buffers must fit TMEM");

  extern __shared__ uint8_t smem_raw[];
  // identical offsets in both CTAs: the dynamic smem base is the same for every CTA of a launch
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared space
  // staging first (1024-byte aligned per-warp 4 KB buffers: the TMA-store drain needs swizzle-atom alignment), then
  // the barriers
  float* stg_all = reinterpret_cast<float*>(smem + STAGES * L::STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * L::STAGE_BYTES + L::STG_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tfull_bar = bars + 2 * STAGES;
  uint64_t* tempty_bar = bars + 2 * STAGES + NBUF;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 2 * NBUF);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = (rank == 0);

  const int m_tiles = (p.M + 2 * BM - 1) / (2 * BM);
  const int n_tiles = (p.N + BN - 1) / BN;
  const int num_kb = (p.K + BK - 1) / BK;
  const int total_work = m_tiles * n_tiles * p.splits;
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  int w_begin, w_end, w_step;
  if (p.raster == 0) { w_begin = pair; w_end = total_work; w_step = num_pairs; }
  else {
    const int per = (total_work + num_pairs - 1) / num_pairs;
    w_begin = pair * per; w_end = min(total_work, w_begin + per); w_step = 1;
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if constexpr (DIRECT == 2) tma_prefetch_desc(&tmap_c);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
      for (int b = 0; b < NBUF; ++b) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], 2 * EPI_WARPS); }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();      // barrier inits and the TMEM allocation of both CTAs are visible pair-wide
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ======================= TMA producer (both CTAs) =======================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = w_begin; w < w_end; w += w_step) {
        const int split = w % p.splits;
        const int t = w / p.splits;
        const int mt = p.raster ? t / n_tiles : t % m_tiles, nt = p.raster ? t % n_tiles : t / m_tiles;
        const int m0 = mt * 2 * BM + (int)rank * BM;
        const int n0 = nt * BN + (int)rank * HB;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(num_kb, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * L::STAGE_BYTES);
          const uint32_t lbar = mapa_rank(smem_u32(&full_bar[stage]), 0);
          uint8_t* sa = smem + stage * L::STAGE_BYTES;
          uint8_t* sb = sa + L::A_BYTES;
          if (A_MN) {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j)
              tma_load_2d_pair(sa + j * (64 * BK * 2), &tmap_a, lbar, m0 + j * 64, kb * BK);
          } else {
            tma_load_2d_pair(sa, &tmap_a, lbar, kb * BK, m0);
          }
          if (B_MN) {
#pragma unroll
            for (int j = 0; j < HB / 64; ++j)
              tma_load_2d_pair(sb + j * (64 * BK * 2), &tmap_b, lbar, n0 + j * 64, kb * BK);
          } else {
            tma_load_2d_pair(sb, &tmap_b, lbar, kb * BK, n0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
      // tail: do not leave while the leader's multicast commits can still arrive on this CTA's empty barriers
      for (int s = 0; s < STAGES; ++s) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer (leader CTA only) ===================
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc_m(2 * BM, BN, A_MN, B_MN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int w = w_begin; w < w_end; w += w_step, ++it) {
        const int split = w % p.splits;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(num_kb, kb0 + p.kb_per_split);
        const int buf = it & (NBUF - 1);
        const uint32_t acc_phase = (uint32_t)(it / NBUF) & 1u;
        mbar_wait(&tempty_bar[buf], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(buf * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * L::STAGE_BYTES);
          const uint32_t sb = sa + L::A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t ad = A_MN ? make_sdesc(sa + k * 2048, 64 * BK * 2, 1024) : make_sdesc(sa + k * 32, 16, 1024);
            const uint64_t bd = B_MN ? make_sdesc(sb + k * 2048, 64 * BK * 2, 1024) : make_sdesc(sb + k * 32, 16, 1024);
            if (!(p.dbg & 2)) umma_bf16_pair(tmem_d, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit_pair(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_pair(&tfull_bar[buf]);
      }
    }
  } else {
    // ======================= epilogue warps (both CTAs) =====================
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    float* stg = stg_all + (size_t)(warp - 2) * 32 * STG_PITCH;
    int it = 0;
    for (int w = w_begin; w < w_end; w += w_step, ++it) {
      const int split = w % p.splits;
      const int t = w / p.splits;
      const int mt = p.raster ? t / n_tiles : t % m_tiles, nt = p.raster ? t % n_tiles : t / m_tiles;
      const int buf = it & (NBUF - 1);
      const uint32_t acc_phase = (uint32_t)(it / NBUF) & 1u;
      mbar_wait(&tfull_bar[buf], acc_phase);
      tc_fence_after();
      const int m_base = mt * 2 * BM + (int)rank * BM + quad * 32;
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * BN);
      OutSel o{p.C, p.c_f32, p.mode, p.epi, p.beta};
      if (p.splits > 1) {
        o.C = p.ws + (long long)split * p.ws_slab; o.c_f32 = 1; o.mode = EM_NONE; o.epi = LR2_EPI_NONE; o.beta = 0.f;
      }
      if constexpr (DIRECT == 2) {
        static_assert(DIRECT != 2 || (BN == 256 && EPI_WARPS == 16), "TMA-store drain: 64 columns per warp");
        if (!(p.dbg & 1) && m_base < p.M && nt * BN + half * 64 < p.N)      // warp-uniform: box not entirely outside
          drain_warp_tma(&tmap_c, reinterpret_cast<uint8_t*>(stg), taddr, half * 64, lane, m_base, nt * BN + half * 64);
      } else {
        if (!(p.dbg & 1)) drain_tile<BN, false, DIRECT>(p, o, taddr, m_base, nt, stg, lane, half);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_rank(smem_u32(&tempty_bar[buf]), 0));
    }
    if constexpr (DIRECT == 2) {
      if (lane == 0) tma_store_wait_read();    // the staging buffer must outlive the last box read
    }
  }

  __syncwarp();
  tc_fence_before();
  cluster_sync_exec();     // neither CTA may exit (or free TMEM) while its pair still reads its smem / TMEM
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// Fold split-K slabs and apply the epilogue. One thread per 8 output columns.
__global__ void splitk_reduce_kernel(const __grid_constant__ GemmParams p, long long out_rows, int out_cols) {
  const OutSel o{p.C, p.c_f32, p.mode, p.epi, p.beta};
  const int cols8 = out_cols / 8;
  const long long total = out_rows * cols8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols8;
    const int c = (int)(i % cols8) * 8;
    float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int s = 0; s < p.splits; ++s) {
      const float4* src = reinterpret_cast<const float4*>(p.ws + (long long)s * p.ws_slab + r * p.ldc + c);
      const float4 a = src[0], b = src[1];
      v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w;
      v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
    }
    epi_store8(p, o, v, r, c);
  }
}

// ------------------------------------------------------------ host side --
static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
static std::mutex g_mu;

static int ensure_encode() {
  if (g_encode) return LR2_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || fn == nullptr) return LR2_ERR_TMA;
  g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  return LR2_OK;
}

struct MapKey {
  const void* ptr; long long d0, d1, stride; int b0, b1;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && d0 == o.d0 && d1 == o.d1 && stride == o.stride && b0 == o.b0 && b1 == o.b1;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    h = h * 1000003u ^ (size_t)k.d0; h = h * 1000003u ^ (size_t)k.d1; h = h * 1000003u ^ (size_t)k.stride;
    h = h * 1000003u ^ (size_t)(k.b0 * 1024 + k.b1);
    return h;
  }
};
static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;

// 2-D bf16 tensor map: inner dim d0 (contiguous), outer dim d1 with row pitch `stride` elements;
// box = b0 x b1, 128-byte swizzle, OOB -> zero fill.
static int get_tmap(const void* ptr, long long d0, long long d1, long long stride, int b0, int b1, CUtensorMap* out) {
  std::lock_guard<std::mutex> lk(g_mu);
  int rc = ensure_encode();
  if (rc != LR2_OK) return rc;
  MapKey key{ptr, d0, d1, stride, b0, b1};
  auto it = g_maps.find(key);
  if (it != g_maps.end()) { *out = it->second; return LR2_OK; }
  cuuint64_t gdim[2] = {(cuuint64_t)d0, (cuuint64_t)d1};
  cuuint64_t gstr[1] = {(cuuint64_t)stride * 2};
  cuuint32_t box[2] = {(cuuint32_t)b0, (cuuint32_t)b1};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return LR2_ERR_TMA;
  if (g_maps.size() > 4096) g_maps.clear();
  g_maps.emplace(key, m);
  *out = m;
  return LR2_OK;
}

static int g_num_sms = 0;
static int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

template <int BN, bool A_MN, bool B_MN, bool ADAMW = false>
static int launch(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t stream) {
  using L = SmemLayout<BN>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_kernel<BN, A_MN, B_MN, ADAMW>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
    if (e != cudaSuccess) return LR2_ERR_CUDA;
    configured = true;
  }
  const int m_tiles = (p.M + BM - 1) / BM, n_tiles = (p.N + BN - 1) / BN;
  const int total = m_tiles * n_tiles * p.splits;
  const int grid = total < num_sms() ? total : num_sms();
  gemm_kernel<BN, A_MN, B_MN, ADAMW><<<grid, GEMM_THREADS, L::TOTAL, stream>>>(ta, tb, p); LR2_LAUNCHED(1);
  LR2_RETURN_LAUNCH();
}

// 2-CTA pair kernel: persistent over min(#tiles, resident clusters) CTA pairs.
template <int BN, bool A_MN, bool B_MN, int DIRECT>
static int launch2(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const GemmParams& p,
                   cudaStream_t stream) {
  using L = SmemLayout2<BN>;
  static int max_pairs = 0;
  if (max_pairs == 0) {
    cudaError_t e = cudaFuncSetAttribute(gemm2_kernel<BN, A_MN, B_MN, DIRECT>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
    if (e != cudaSuccess) return LR2_ERR_CUDA;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(num_sms() & ~1); cfg.blockDim = dim3(GEMM_THREADS); cfg.dynamicSmemBytes = L::TOTAL;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    e = cudaOccupancyMaxActiveClusters(&n, gemm2_kernel<BN, A_MN, B_MN, DIRECT>, &cfg);
    if (e != cudaSuccess || n <= 0) { cudaGetLastError(); n = num_sms() / 2; }
    max_pairs = n < num_sms() / 2 ? n : num_sms() / 2;
  }
  const int m_tiles = (p.M + 2 * BM - 1) / (2 * BM), n_tiles = (p.N + BN - 1) / BN;
  const int total = m_tiles * n_tiles * p.splits;
  const int pairs = total < max_pairs ? total : max_pairs;
  gemm2_kernel<BN, A_MN, B_MN, DIRECT><<<2 * pairs, GEMM_THREADS, L::TOTAL, stream>>>(ta, tb, tc, p); LR2_LAUNCHED(1);
  LR2_RETURN_LAUNCH();
}

template <int BN, int DIRECT>
static int launch2_major(bool a_mn, bool b_mn, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc,
                         const GemmParams& p, cudaStream_t s) {
  if (!a_mn && !b_mn) return launch2<BN, false, false, DIRECT>(ta, tb, tc, p, s);
  if (!a_mn && b_mn) return launch2<BN, false, true, DIRECT>(ta, tb, tc, p, s);
  if (a_mn && !b_mn) return launch2<BN, true, false, DIRECT>(ta, tb, tc, p, s);
  return launch2<BN, true, true, DIRECT>(ta, tb, tc, p, s);
}

template <int BN>
static int launch_major(bool a_mn, bool b_mn, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p,
                        cudaStream_t s) {
  if (!a_mn && !b_mn) return launch<BN, false, false>(ta, tb, p, s);
  if (!a_mn && b_mn) return launch<BN, false, true>(ta, tb, p, s);
  if (a_mn && !b_mn) return launch<BN, true, false>(ta, tb, p, s);
  return launch<BN, true, true>(ta, tb, p, s);
}

}  // namespace lr2

using namespace lr2;

extern "C" long long lr2_gemm_workspace_bytes(int M, int N, int splits, int transposed_out, long long ldc) {
  if (splits <= 1) return 0;
  const long long rows = transposed_out ? N : M;
  return (long long)splits * rows * ldc * 4;
}

extern "C" int lr2_gemm_bf16(const void* A, long long lda, int a_mn_major, const void* B, long long ldb,
                             int b_mn_major, void* C, long long ldc, int c_is_f32, int transposed_out, int M, int N,
                             int K, int epilogue, const float* bias, const void* aux, long long ldaux, void* C2,
                             float beta, float drop_p, unsigned long long seed, unsigned int site,
                             const void* seed_dev, int splits,
                             void* workspace, int block_n, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (M <= 0 || N <= 0 || K <= 0) return LR2_ERR_BAD_SHAPE;
  if (epilogue < 0 || epilogue > LR2_EPI_DQGELU || epilogue == LR2_EPI_ADAMW) return LR2_ERR_UNSUPPORTED;
  int act = 0;
  if (epilogue == LR2_EPI_BIAS_QGELU) { epilogue = LR2_EPI_BIAS_GELU; act = 1; }
  if (epilogue == LR2_EPI_DQGELU) { epilogue = LR2_EPI_DGELU; act = 1; }
  if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B) | reinterpret_cast<uintptr_t>(C)) & 15)
    return LR2_ERR_MISALIGNED;
  if ((lda % 8) || (ldb % 8)) return LR2_ERR_MISALIGNED;
  const int out_cols = transposed_out ? M : N;
  if (!transposed_out && ((ldc % 8) || (N % 8))) return LR2_ERR_MISALIGNED;
  if (transposed_out && drop_p > 0.f) return LR2_ERR_UNSUPPORTED;
  if (splits < 1) splits = 1;
  const int num_kb = (K + BK - 1) / BK;
  if (splits > num_kb) splits = num_kb;
  int kb_per = (num_kb + splits - 1) / splits;
  splits = (num_kb + kb_per - 1) / kb_per;  // no empty splits
  if (splits > 1) {
    if (workspace == nullptr) return LR2_ERR_BAD_SHAPE;
    if ((ldc % 8) || (out_cols % 8)) return LR2_ERR_MISALIGNED;
  }
  int BN = block_n;
  bool pair = false;          // block_n = 2000 + BN selects the cta_group::2 pair kernel (256 x BN tiles)
  if (BN == 0) {
    static int auto_pair = -1;
    if (auto_pair < 0) { const char* e = getenv("LR2_GEMM_PAIR"); auto_pair = e ? atoi(e) : 1; }
    // untransposed problems whose 256 x 256 pair tiles keep at least half of the 74 CTA pairs busy run on the
    // cta_group::2 kernel (half the L2->SMEM bytes per FLOP); everything else on the single-CTA kernel.  Callers that
    // want split-K to fill the pairs pick `splits` with the same rule (lr2ppo_b200/ops.py: plan_gemm).
    const long long t256 = (long long)((M + 255) / 256) * ((N + 255) / 256) * splits;
    if (auto_pair && !transposed_out && N % 256 == 0 && M >= 256 && K > 128 && t256 >= 37) {
      pair = true;
      BN = 256;
    } else {
      BN = (N <= 64) ? 64 : (N <= 128 ? 128 : ((N > 128 && N <= 256 && transposed_out) ? 256 : 128));
    }
  } else if (BN >= 2000) {
    pair = true; BN -= 2000;
    if (BN != 128 && BN != 256) return LR2_ERR_UNSUPPORTED;
  }
  if (BN != 64 && BN != 128 && BN != 256) return LR2_ERR_UNSUPPORTED;

  CUtensorMap ta, tb;
  int rc;
  // K-major operand [rows,K]: dims {K, rows}, box {64, tile_rows}; MN-major [K,rows]: dims {rows, K}, box {64, 64}.
  rc = a_mn_major ? get_tmap(A, M, K, lda, 64, BK, &ta) : get_tmap(A, K, M, lda, BK, BM, &ta);
  if (rc != LR2_OK) return rc;
  rc = b_mn_major ? get_tmap(B, N, K, ldb, 64, BK, &tb) : get_tmap(B, K, N, ldb, BK, pair ? BN / 2 : BN, &tb);
  if (rc != LR2_OK) return rc;

  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.N = N; p.K = K;
  p.splits = splits; p.kb_per_split = kb_per;
  p.epi = epilogue; p.transposed_out = transposed_out; p.c_f32 = c_is_f32; p.act = act;
  p.C = C; p.ldc = ldc; p.C2 = reinterpret_cast<bf16*>(C2);
  p.bias = bias; p.aux = reinterpret_cast<const bf16*>(aux); p.ldaux = ldaux;
  p.beta = beta; p.drop_p = drop_p; p.seed = seed; p.site = site;
  p.seed_dev = reinterpret_cast<const unsigned long long*>(seed_dev);
  p.drop_thresh = dropout_thresh16(drop_p); p.drop_scale = dropout_scale16(drop_p);
  p.mode = resolve_mode(epilogue, act, drop_p, bias, c_is_f32, beta);
  { static int d = -1; if (d < 0) { const char* e = getenv("LR2_GEMM_DBG"); d = e ? atoi(e) : 0; } p.dbg = d; }
  CUtensorMap tc = ta;        // only read by the TMA-store instantiation
  {
    // Plain bf16 outputs of the pair kernel (256-wide tiles, no split-K) skip the generic staged epilogue:
    //   LR2_GEMM_DIRECT=2 (default)  TMA-store drain: {64, 32} boxes of an output tensor map (any 16-byte aligned pitch)
    //   LR2_GEMM_DIRECT=1            256-bit register stores (needs 32-byte aligned row pieces)
    //   LR2_GEMM_DIRECT=0            staged drain, as every other epilogue mode
    static int dsel = -1;
    if (dsel < 0) { const char* e = getenv("LR2_GEMM_DIRECT"); dsel = e ? atoi(e) : 2; }
    const bool plain = pair && BN == 256 && p.mode == EM_NONE && !transposed_out && !c_is_f32 && splits == 1;
    p.direct = 0;
    if (plain && dsel >= 2) {
      rc = get_tmap(C, N, M, ldc, 64, 32, &tc);
      if (rc != LR2_OK) return rc;
      p.direct = 2;
    } else if (plain && dsel == 1 && (ldc % 16 == 0) && ((reinterpret_cast<uintptr_t>(C) & 31) == 0)) {
      p.direct = 1;
    }
  }
  p.ws = reinterpret_cast<float*>(workspace);
  { static int r = -1; if (r < 0) { const char* e = getenv("LR2_GEMM_RASTER"); r = e ? atoi(e) : 0; } p.raster = r; }
  const long long out_rows = transposed_out ? N : M;
  p.ws_slab = out_rows * ldc;

  const bool amn = a_mn_major != 0, bmn = b_mn_major != 0;
  if (pair && BN == 256 && p.direct == 2) rc = launch2_major<256, 2>(amn, bmn, ta, tb, tc, p, stream);
  else if (pair && BN == 256 && p.direct == 1) rc = launch2_major<256, 1>(amn, bmn, ta, tb, tc, p, stream);
  else if (pair && BN == 256) rc = launch2_major<256, 0>(amn, bmn, ta, tb, tc, p, stream);
  else if (pair) rc = launch2_major<128, 0>(amn, bmn, ta, tb, tc, p, stream);
  else if (BN == 64) rc = launch_major<64>(a_mn_major != 0, b_mn_major != 0, ta, tb, p, stream);
  else if (BN == 128) rc = launch_major<128>(a_mn_major != 0, b_mn_major != 0, ta, tb, p, stream);
  else rc = launch_major<256>(a_mn_major != 0, b_mn_major != 0, ta, tb, p, stream);
  if (rc != LR2_OK) return rc;

  if (splits > 1) {
    const long long total = out_rows * (out_cols / 8);
    int blocks = (int)((total + 255) / 256);
    const int cap = num_sms() * 8;
    if (blocks > cap) blocks = cap;
    splitk_reduce_kernel<<<blocks, 256, 0, stream>>>(p, out_rows, out_cols); LR2_LAUNCHED(1);
    LR2_RETURN_LAUNCH();
  }
  return LR2_OK;
}

int lr2_adamw_wgrad_mma(const void* dY, long long lddy, const void* X, long long ldx, int rows, int out_f, int in_f,
                        float* param, float* exp_avg, float* exp_avg_sq, void* shadow_bf16, const float* hyper,
                        float weight_decay, cudaStream_t stream);   // adamw_wgrad.cu

// Fused out_layer.fc1 weight gradient + AdamW: the [out, in] fp32 parameter is updated tile by tile from the
// TMEM accumulator of dY^T X, so the 500 M-element gradient is never written to or read from HBM.
// HBM bytes per parameter: read p, m, v (12) + write p, m, v (12) + bf16 shadow (2) = 26 (28+2 unfused, plus
// the 4+4 of writing and re-reading the gradient).
extern "C" int lr2_gemm_wgrad_adamw(const void* dY, long long lddy, const void* X, long long ldx, int rows, int out_f,
                                    int in_f, float* param, float* exp_avg, float* exp_avg_sq, void* shadow_bf16,
                                    const float* hyper, float weight_decay, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (rows <= 0 || out_f <= 0 || in_f <= 0) return LR2_ERR_BAD_SHAPE;
  if ((lddy % 8) || (ldx % 8) || (in_f % 8)) return LR2_ERR_MISALIGNED;
  if ((reinterpret_cast<uintptr_t>(dY) | reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(param) |
       reinterpret_cast<uintptr_t>(exp_avg) | reinterpret_cast<uintptr_t>(exp_avg_sq) |
       reinterpret_cast<uintptr_t>(shadow_bf16)) & 15)
    return LR2_ERR_MISALIGNED;
  if (hyper == nullptr) return LR2_ERR_BAD_SHAPE;
  {
    // opt-in linear-pass implementation for short reductions (adamw_wgrad.cu; not GPU-validated yet)
    static const int impl_mma = [] { const char* e = getenv("LR2_WGRAD_ADAMW_IMPL"); return e && !strcmp(e, "mma") ? 1 : 0; }();
    if (impl_mma) {
      const int rc_mma = lr2_adamw_wgrad_mma(dY, lddy, X, ldx, rows, out_f, in_f, param, exp_avg, exp_avg_sq, shadow_bf16,
                                             hyper, weight_decay, stream);
      if (rc_mma != LR2_ERR_UNSUPPORTED) return rc_mma;
    }
  }
  static int BN_sel = 0;
  if (BN_sel == 0) { const char* e = getenv("LR2_ADAMW_BN"); BN_sel = e ? atoi(e) : 128; }
  CUtensorMap ta, tb;
  int rc = get_tmap(dY, out_f, rows, lddy, 64, BK, &ta);   // A: MN-major [K=rows, M=out_f]
  if (rc != LR2_OK) return rc;
  rc = get_tmap(X, in_f, rows, ldx, 64, BK, &tb);           // B: MN-major [K=rows, N=in_f]
  if (rc != LR2_OK) return rc;
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = out_f; p.N = in_f; p.K = rows;
  p.splits = 1; p.kb_per_split = (rows + BK - 1) / BK;
  p.epi = LR2_EPI_ADAMW; p.transposed_out = 0; p.c_f32 = 1;
  p.C = param; p.ldc = in_f;
  p.drop_scale = 1.f;
  p.raster = 1;   // each CTA walks along a row panel: long contiguous p/m/v streams per row
  p.adam_m = exp_avg; p.adam_v = exp_avg_sq; p.adam_shadow = reinterpret_cast<bf16*>(shadow_bf16);
  p.adam_hyper = hyper; p.adam_wd = weight_decay;
  if (BN_sel == 256) return launch<256, true, true, true>(ta, tb, p, stream);
  return launch<128, true, true, true>(ta, tb, p, stream);
}
