// Multi-headed self-attention core of the TencentPretrain towers (ViT-B/16: S = 197, RoBERTa-base: S <= 256;
// 12 heads x 64):  P = softmax(Q K^T * scale + key_bias),  O = dropout(P) V
// ref: tencentpretrain/layers/multi_headed_attn.py:55-76 (scale 1/sqrt(d_head) BEFORE softmax, additive -10000
// mask built from seg in encoders/transformer_encoder.py:62-68, dropout on the probabilities).
// Flash-style: the S x S score matrix never leaves the SM.  One CTA per (batch, head); K and V of the head live in
// shared memory; TWO threads own one query row (32 of the 64 head dims each, one shuffle per dot product) and
// run an online softmax.  Backward recomputes P from the saved log-sum-exp: phase A (per query row) produces dQ
// and D = rowsum(dO*O), phase B (per key row) produces dK and dV.  fp32 math, bf16 I/O.
// 4*S*S*64 FLOP per head against 4*S*64*2 bytes: compute-bound on CUDA cores for S ~ 200; a tcgen05 version is
// listed as next work in DESIGN.md.
#include <cstdlib>
#include "common.cuh"

namespace lr2 {

constexpr int MHA_DH = 64;
constexpr int MHA_HALF = 32;
constexpr int MHA_MAX_S = 256;
constexpr int MHA_THREADS = 512;  // 2 threads per row

__device__ __forceinline__ void ld8b(const bf16* p, float* v) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 t;
  t = unpack_bf16x2(u.x); v[0] = t.x; v[1] = t.y;
  t = unpack_bf16x2(u.y); v[2] = t.x; v[3] = t.y;
  t = unpack_bf16x2(u.z); v[4] = t.x; v[5] = t.y;
  t = unpack_bf16x2(u.w); v[6] = t.x; v[7] = t.y;
}
__device__ __forceinline__ void st8b(bf16* p, const float* v) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
  u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}

// keep bit of probability (row i, key j) of head bh; bits for 4 consecutive j come from one Philox call
__device__ __forceinline__ uint32_t mha_keep4(unsigned long long seed, long long bh, int i, int j4, int S4,
                                              uint32_t thresh) {
  const unsigned long long idx4 = ((unsigned long long)bh * (unsigned long long)MHA_MAX_S + (unsigned long long)i) *
                                      (unsigned long long)(S4 / 4) + (unsigned long long)(j4 / 4);
  return dropout_keep4(seed, 0x4D48u, idx4, thresh);
}

__global__ void __launch_bounds__(MHA_THREADS)
mha_fwd_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k, const bf16* __restrict__ v, long long ld,
               const float* __restrict__ key_bias, bf16* __restrict__ o, long long ldo, float* __restrict__ lse,
               int S, int H, float scale, float drop_p, uint32_t thresh, unsigned long long seed_in,
               const unsigned long long* __restrict__ seed_dev) {
  extern __shared__ __align__(16) float msm[];
  float* Ks = msm;                    // [S][64]
  float* Vs = Ks + (size_t)S * MHA_DH;
  float* Bs = Vs + (size_t)S * MHA_DH;  // [S] additive key bias
  const int bh = blockIdx.x, b = bh / H, h = bh % H;
  const unsigned long long seed = seed_in + (seed_dev ? *seed_dev : 0ull);
  const long long row0 = (long long)b * S;
  for (int i = threadIdx.x; i < S * (MHA_DH / 8); i += blockDim.x) {
    const int r = i / (MHA_DH / 8), c = (i % (MHA_DH / 8)) * 8;
    ld8b(k + (row0 + r) * ld + h * MHA_DH + c, Ks + r * MHA_DH + c);
    ld8b(v + (row0 + r) * ld + h * MHA_DH + c, Vs + r * MHA_DH + c);
  }
  for (int j = threadIdx.x; j < S; j += blockDim.x) Bs[j] = key_bias ? key_bias[(long long)b * S + j] : 0.f;
  __syncthreads();
  const int row_raw = threadIdx.x >> 1, half = threadIdx.x & 1;
  const bool active = row_raw < S;
  const int row = active ? row_raw : S - 1;  // idle lanes shadow the last row so warp shuffles stay full-mask
  const int d0 = half * MHA_HALF;
  float qv[MHA_HALF], acc[MHA_HALF];
#pragma unroll
  for (int c = 0; c < MHA_HALF; c += 8) ld8b(q + (row0 + row) * ld + h * MHA_DH + d0 + c, qv + c);
#pragma unroll
  for (int c = 0; c < MHA_HALF; ++c) acc[c] = 0.f;
  float m = -INFINITY, l = 0.f;
  const int S4 = (S + 3) & ~3;
  const float dscale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  uint32_t keep = 0xFu;
  for (int j = 0; j < S; ++j) {
    const float4* kr = reinterpret_cast<const float4*>(Ks + j * MHA_DH + d0);
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < MHA_HALF / 4; ++c) {
      const float4 kk = kr[c];
      s += qv[4 * c] * kk.x + qv[4 * c + 1] * kk.y + qv[4 * c + 2] * kk.z + qv[4 * c + 3] * kk.w;
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s = s * scale + Bs[j];
    if (s > m) {
      const float corr = __expf(m - s);
      l *= corr;
#pragma unroll
      for (int c = 0; c < MHA_HALF; ++c) acc[c] *= corr;
      m = s;
    }
    const float p = __expf(s - m);
    l += p;
    float pd = p;
    if (drop_p > 0.f) {
      if ((j & 3) == 0) keep = mha_keep4(seed, bh, row, j, S4, thresh);
      pd = ((keep >> (j & 3)) & 1u) ? p * dscale : 0.f;
    }
    const float4* vr = reinterpret_cast<const float4*>(Vs + j * MHA_DH + d0);
#pragma unroll
    for (int c = 0; c < MHA_HALF / 4; ++c) {
      const float4 vv = vr[c];
      acc[4 * c] += pd * vv.x; acc[4 * c + 1] += pd * vv.y; acc[4 * c + 2] += pd * vv.z; acc[4 * c + 3] += pd * vv.w;
    }
  }
  const float inv = 1.f / l;
#pragma unroll
  for (int c = 0; c < MHA_HALF; ++c) acc[c] *= inv;
  if (active) {
#pragma unroll
    for (int c = 0; c < MHA_HALF; c += 8) st8b(o + (row0 + row) * ldo + h * MHA_DH + d0 + c, acc + c);
    if (lse != nullptr && half == 0) lse[(long long)bh * S + row] = m + __logf(l);
  }
}

__global__ void __launch_bounds__(MHA_THREADS)
mha_bwd_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k, const bf16* __restrict__ v, long long ld,
               const float* __restrict__ key_bias, const bf16* __restrict__ o, const bf16* __restrict__ d_o,
               long long ldo, const float* __restrict__ lse, bf16* __restrict__ dq, bf16* __restrict__ dk,
               bf16* __restrict__ dv, long long ldd, int S, int H, float scale, float drop_p, uint32_t thresh,
               unsigned long long seed_in, const unsigned long long* __restrict__ seed_dev) {
  extern __shared__ __align__(16) unsigned char braw[];
  bf16* Qs = reinterpret_cast<bf16*>(braw);      // [S][64] each
  bf16* Ks = Qs + (size_t)S * MHA_DH;
  bf16* Vs = Ks + (size_t)S * MHA_DH;
  bf16* Gs = Vs + (size_t)S * MHA_DH;            // dO
  float* Ls = reinterpret_cast<float*>(Gs + (size_t)S * MHA_DH);  // lse [S]
  float* Ds = Ls + S;                            // D = rowsum(dO * O) [S]
  float* Bs = Ds + S;                            // key bias [S]
  const int bh = blockIdx.x, b = bh / H, h = bh % H;
  const unsigned long long seed = seed_in + (seed_dev ? *seed_dev : 0ull);
  const long long row0 = (long long)b * S;
  for (int i = threadIdx.x; i < S * (MHA_DH / 8); i += blockDim.x) {
    const int r = i / (MHA_DH / 8), c = (i % (MHA_DH / 8)) * 8;
    const long long g = (row0 + r) * ld + h * MHA_DH + c;
    *reinterpret_cast<uint4*>(Qs + r * MHA_DH + c) = *reinterpret_cast<const uint4*>(q + g);
    *reinterpret_cast<uint4*>(Ks + r * MHA_DH + c) = *reinterpret_cast<const uint4*>(k + g);
    *reinterpret_cast<uint4*>(Vs + r * MHA_DH + c) = *reinterpret_cast<const uint4*>(v + g);
    *reinterpret_cast<uint4*>(Gs + r * MHA_DH + c) =
        *reinterpret_cast<const uint4*>(d_o + (row0 + r) * ldo + h * MHA_DH + c);
  }
  for (int j = threadIdx.x; j < S; j += blockDim.x) {
    Ls[j] = lse[(long long)bh * S + j];
    Bs[j] = key_bias ? key_bias[(long long)b * S + j] : 0.f;
  }
  __syncthreads();
  const int row_raw = threadIdx.x >> 1, half = threadIdx.x & 1;
  const bool active = row_raw < S;
  const int row = active ? row_raw : S - 1;  // idle lanes shadow the last row (full-mask shuffles), stores predicated
  const int d0 = half * MHA_HALF;
  const int S4 = (S + 3) & ~3;
  const float dscale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;

  // ---- phase A: per query row i: D_i = rowsum(dO_i * O_i) and dQ_i = sum_j dS_ij k_j * scale ----
  {
    float xq[MHA_HALF], xg[MHA_HALF], dqa[MHA_HALF];
    float dsum = 0.f;
#pragma unroll
    for (int c = 0; c < MHA_HALF; c += 8) {
      float ov[8];
      ld8b(Qs + row * MHA_DH + d0 + c, xq + c);
      ld8b(Gs + row * MHA_DH + d0 + c, xg + c);
      ld8b(o + (row0 + row) * ldo + h * MHA_DH + d0 + c, ov);
#pragma unroll
      for (int e = 0; e < 8; ++e) { dsum += xg[c + e] * ov[e]; dqa[c + e] = 0.f; }
    }
    dsum += __shfl_xor_sync(0xffffffffu, dsum, 1);
    if (active && half == 0) Ds[row] = dsum;
    const float li = Ls[row];
    uint32_t keep = 0xFu;
    for (int j = 0; j < S; ++j) {
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int c = 0; c < MHA_HALF; c += 8) {
        float kk[8], vv[8];
        ld8b(Ks + j * MHA_DH + d0 + c, kk);
        ld8b(Vs + j * MHA_DH + d0 + c, vv);
#pragma unroll
        for (int e = 0; e < 8; ++e) { s += xq[c + e] * kk[e]; dp += xg[c + e] * vv[e]; }
      }
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      dp += __shfl_xor_sync(0xffffffffu, dp, 1);
      const float p = __expf(s * scale + Bs[j] - li);
      if (drop_p > 0.f) {
        if ((j & 3) == 0) keep = mha_keep4(seed, bh, row, j, S4, thresh);
        dp = ((keep >> (j & 3)) & 1u) ? dp * dscale : 0.f;
      }
      const float ds = p * (dp - dsum) * scale;
#pragma unroll
      for (int c = 0; c < MHA_HALF; c += 8) {
        float kk[8];
        ld8b(Ks + j * MHA_DH + d0 + c, kk);
#pragma unroll
        for (int e = 0; e < 8; ++e) dqa[c + e] += ds * kk[e];
      }
    }
    if (active) {
#pragma unroll
      for (int c = 0; c < MHA_HALF; c += 8) st8b(dq + (row0 + row) * ldd + h * MHA_DH + d0 + c, dqa + c);
    }
  }
  __syncthreads();

  // ---- phase B1: per key row j: dV_j = sum_i Pdrop_ij dO_i ----
  const int j = row;
  const float bj = Bs[j];
  {
    float kk[MHA_HALF], dva[MHA_HALF];
#pragma unroll
    for (int c = 0; c < MHA_HALF; c += 8) ld8b(Ks + j * MHA_DH + d0 + c, kk + c);
#pragma unroll
    for (int c = 0; c < MHA_HALF; ++c) dva[c] = 0.f;
    for (int i = 0; i < S; ++i) {
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < MHA_HALF; c += 8) {
        float qi[8];
        ld8b(Qs + i * MHA_DH + d0 + c, qi);
#pragma unroll
        for (int e = 0; e < 8; ++e) s += qi[e] * kk[c + e];
      }
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      float pd = __expf(s * scale + bj - Ls[i]);
      if (drop_p > 0.f) {
        const uint32_t kb = mha_keep4(seed, bh, i, j & ~3, S4, thresh);
        pd = ((kb >> (j & 3)) & 1u) ? pd * dscale : 0.f;
      }
#pragma unroll
      for (int c = 0; c < MHA_HALF; c += 8) {
        float gi[8];
        ld8b(Gs + i * MHA_DH + d0 + c, gi);
#pragma unroll
        for (int e = 0; e < 8; ++e) dva[c + e] += pd * gi[e];
      }
    }
    if (active) {
#pragma unroll
      for (int c = 0; c < MHA_HALF; c += 8) st8b(dv + (row0 + j) * ldd + h * MHA_DH + d0 + c, dva + c);
    }
  }
  // ---- phase B2: per key row j: dK_j = sum_i dS_ij q_i * scale ----
  {
    float kk[MHA_HALF], vv[MHA_HALF], dka[MHA_HALF];
#pragma unroll
    for (int c = 0; c < MHA_HALF; c += 8) {
      ld8b(Ks + j * MHA_DH + d0 + c, kk + c);
      ld8b(Vs + j * MHA_DH + d0 + c, vv + c);
    }
#pragma unroll
    for (int c = 0; c < MHA_HALF; ++c) dka[c] = 0.f;
    for (int i = 0; i < S; ++i) {
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int c = 0; c < MHA_HALF; c += 8) {
        float qi[8], gi[8];
        ld8b(Qs + i * MHA_DH + d0 + c, qi);
        ld8b(Gs + i * MHA_DH + d0 + c, gi);
#pragma unroll
        for (int e = 0; e < 8; ++e) { s += qi[e] * kk[c + e]; dp += gi[e] * vv[c + e]; }
      }
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      dp += __shfl_xor_sync(0xffffffffu, dp, 1);
      const float p = __expf(s * scale + bj - Ls[i]);
      if (drop_p > 0.f) {
        const uint32_t kb = mha_keep4(seed, bh, i, j & ~3, S4, thresh);
        dp = ((kb >> (j & 3)) & 1u) ? dp * dscale : 0.f;
      }
      const float ds = p * (dp - Ds[i]) * scale;
#pragma unroll
      for (int c = 0; c < MHA_HALF; c += 8) {
        float qi[8];
        ld8b(Qs + i * MHA_DH + d0 + c, qi);
#pragma unroll
        for (int e = 0; e < 8; ++e) dka[c + e] += ds * qi[e];
      }
    }
    if (active) {
#pragma unroll
      for (int c = 0; c < MHA_HALF; c += 8) st8b(dk + (row0 + j) * ldd + h * MHA_DH + d0 + c, dka + c);
    }
  }
}

}  // namespace lr2

using namespace lr2;

// tcgen05 kernels (mha_tc.cu); the CUDA-core kernels above remain as the LR2_MHA_LEGACY=1 cross-check
int lr2_mha_tc_fwd(const void* q, const void* k, const void* v, long long ld, const float* key_bias, void* o,
                   long long ldo, float* lse, int B, int S, int H, float scale, float drop_p, unsigned long long seed,
                   const void* seed_dev, cudaStream_t stream);
int lr2_mha_tc_bwd(const void* q, const void* k, const void* v, long long ld, const float* key_bias, const void* o,
                   const void* d_o, long long ldo, const float* lse, void* dq, void* dk, void* dv, long long ldd, int B,
                   int S, int H, float scale, float drop_p, unsigned long long seed, const void* seed_dev,
                   cudaStream_t stream);
static bool mha_legacy() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("LR2_MHA_LEGACY"); v = (e && atoi(e)) ? 1 : 0; }
  return v == 1;
}

static int mha_check(int B, int S, int H, int dh, long long ld, long long ldo) {
  if (B <= 0 || S <= 0 || H <= 0) return LR2_ERR_BAD_SHAPE;
  if (dh != MHA_DH || S > MHA_MAX_S) return LR2_ERR_UNSUPPORTED;
  if ((ld % 8) || (ldo % 8)) return LR2_ERR_MISALIGNED;
  return LR2_OK;
}

extern "C" int lr2_mha_fwd(const void* q, const void* k, const void* v, long long ld, const float* key_bias, void* o,
                           long long ldo, float* lse, int B, int S, int H, int dh, float scale, float drop_p,
                           unsigned long long seed, const void* seed_dev, void* stream) {
  int rc = mha_check(B, S, H, dh, ld, ldo);
  if (rc != LR2_OK) return rc;
  if (!mha_legacy())
    return lr2_mha_tc_fwd(q, k, v, ld, key_bias, o, ldo, lse, B, S, H, scale, drop_p, seed, seed_dev,
                          reinterpret_cast<cudaStream_t>(stream));
  const size_t smem = ((size_t)2 * S * MHA_DH + S) * sizeof(float);
  static size_t configured = 0;
  if (smem > configured) {
    if (cudaFuncSetAttribute(mha_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return LR2_ERR_CUDA;
    configured = smem;
  }
  int threads = ((2 * S + 31) / 32) * 32;
  mha_fwd_kernel<<<B * H, threads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const bf16*>(q), reinterpret_cast<const bf16*>(k), reinterpret_cast<const bf16*>(v), ld,
      key_bias, reinterpret_cast<bf16*>(o), ldo, lse, S, H, scale, drop_p, dropout_thresh(drop_p), seed,
      reinterpret_cast<const unsigned long long*>(seed_dev));
  LR2_LAUNCHED(1);
  LR2_RETURN_LAUNCH();
}

extern "C" int lr2_mha_bwd(const void* q, const void* k, const void* v, long long ld, const float* key_bias,
                           const void* o, const void* d_o, long long ldo, const float* lse, void* dq, void* dk,
                           void* dv, long long ldd, int B, int S, int H, int dh, float scale, float drop_p,
                           unsigned long long seed, const void* seed_dev, void* stream) {
  int rc = mha_check(B, S, H, dh, ld, ldo);
  if (rc != LR2_OK) return rc;
  if ((ldd % 8) || lse == nullptr) return LR2_ERR_MISALIGNED;
  if (!mha_legacy())
    return lr2_mha_tc_bwd(q, k, v, ld, key_bias, o, d_o, ldo, lse, dq, dk, dv, ldd, B, S, H, scale, drop_p, seed,
                          seed_dev, reinterpret_cast<cudaStream_t>(stream));
  const size_t smem = (size_t)4 * S * MHA_DH * 2 + (size_t)3 * S * sizeof(float);
  static size_t configured = 0;
  if (smem > configured) {
    if (cudaFuncSetAttribute(mha_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return LR2_ERR_CUDA;
    configured = smem;
  }
  int threads = ((2 * S + 31) / 32) * 32;
  mha_bwd_kernel<<<B * H, threads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const bf16*>(q), reinterpret_cast<const bf16*>(k), reinterpret_cast<const bf16*>(v), ld,
      key_bias, reinterpret_cast<const bf16*>(o), reinterpret_cast<const bf16*>(d_o), ldo, lse,
      reinterpret_cast<bf16*>(dq), reinterpret_cast<bf16*>(dk), reinterpret_cast<bf16*>(dv), ldd, S, H, scale, drop_p,
      dropout_thresh(drop_p), seed, reinterpret_cast<const unsigned long long*>(seed_dev));
  LR2_LAUNCHED(1);
  LR2_RETURN_LAUNCH();
}
