// XiT cross-attention core (tiny KV): per (item, head)
//   P = softmax(pre_scale * Q K^T),  O = (post_scale * P) V,   Skv <= 16
// ref: finetune/xit.py:125-148 (softmax first, THEN divide by sqrt(emb); the causal mask
// is computed and dropped by the reference, so there is no mask here).
// 4*Sq*Skv*dh FLOP per head against (2*Sq + 2*Skv)*dh*2 bytes: memory-bound, CUDA cores.
// One CTA per (item, head); K_h / V_h live in shared memory as fp32 and are read as
// warp-wide broadcasts; one thread owns one query row.
#include "common.cuh"

namespace lr2 {

constexpr int XA_MAX_KV = 16;
constexpr int XA_MAX_DH = 128;

__device__ __forceinline__ void ld8(const bf16* p, float (&v)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 t;
  t = unpack_bf16x2(u.x); v[0] = t.x; v[1] = t.y;
  t = unpack_bf16x2(u.y); v[2] = t.x; v[3] = t.y;
  t = unpack_bf16x2(u.z); v[4] = t.x; v[5] = t.y;
  t = unpack_bf16x2(u.w); v[6] = t.x; v[7] = t.y;
}
__device__ __forceinline__ void st8(bf16* p, const float (&v)[8]) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
  u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}

__global__ void xattn_fwd_kernel(const bf16* __restrict__ q, long long ldq, const bf16* __restrict__ k,
                                 const bf16* __restrict__ v, long long ldkv, bf16* __restrict__ o, long long ldo,
                                 int Sq, int Skv, int H, int dh, float pre_scale, float post_scale) {
  __shared__ float Ks[XA_MAX_KV * XA_MAX_DH];
  __shared__ float Vs[XA_MAX_KV * XA_MAX_DH];
  const int item = blockIdx.x / H, h = blockIdx.x % H;
  const bf16* kb = k + (long long)item * Skv * ldkv + h * dh;
  const bf16* vb = v + (long long)item * Skv * ldkv + h * dh;
  for (int i = threadIdx.x; i < Skv * dh; i += blockDim.x) {
    const int j = i / dh, d = i % dh;
    Ks[j * dh + d] = __bfloat162float(kb[(long long)j * ldkv + d]);
    Vs[j * dh + d] = __bfloat162float(vb[(long long)j * ldkv + d]);
  }
  __syncthreads();
  for (int r = threadIdx.x; r < Sq; r += blockDim.x) {
    const bf16* qr = q + ((long long)item * Sq + r) * ldq + h * dh;
    float s[XA_MAX_KV];
#pragma unroll
    for (int j = 0; j < XA_MAX_KV; ++j) s[j] = 0.f;
    for (int d = 0; d < dh; d += 8) {
      float qv[8];
      ld8(qr + d, qv);
#pragma unroll
      for (int j = 0; j < XA_MAX_KV; ++j) {
        if (j < Skv) {
          const float* kr = Ks + j * dh + d;
#pragma unroll
          for (int i = 0; i < 8; ++i) s[j] += qv[i] * kr[i];
        }
      }
    }
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < XA_MAX_KV; ++j)
      if (j < Skv) { s[j] *= pre_scale; mx = fmaxf(mx, s[j]); }
    float den = 0.f;
#pragma unroll
    for (int j = 0; j < XA_MAX_KV; ++j)
      if (j < Skv) { s[j] = __expf(s[j] - mx); den += s[j]; }
    const float inv = post_scale / den;
#pragma unroll
    for (int j = 0; j < XA_MAX_KV; ++j) s[j] = (j < Skv) ? s[j] * inv : 0.f;
    bf16* orow = o + ((long long)item * Sq + r) * ldo + h * dh;
    for (int d = 0; d < dh; d += 8) {
      float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int j = 0; j < XA_MAX_KV; ++j) {
        if (j < Skv) {
          const float* vr = Vs + j * dh + d;
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] += s[j] * vr[i];
        }
      }
      st8(orow + d, acc);
    }
  }
}

// Backward.  Phase 1 (thread per query row): recompute P, dP, dS; write dQ; stash
// PP = post*P and dS in shared memory.  Phase 2 (thread per (j, 8 d)): dV = PP^T dO,
// dK = dS^T Q as small dense reductions over the Sq rows held in shared memory.
__global__ void xattn_bwd_kernel(const bf16* __restrict__ q, long long ldq, const bf16* __restrict__ k,
                                 const bf16* __restrict__ v, long long ldkv, const bf16* __restrict__ d_o,
                                 long long ldo, bf16* __restrict__ dq, long long lddq, bf16* __restrict__ dk,
                                 bf16* __restrict__ dv, long long lddkv, int Sq, int Skv, int H, int dh,
                                 float pre_scale, float post_scale) {
  extern __shared__ __align__(16) uint8_t xs[];
  const int dhp = dh + 8;  // padded row (bf16) -> conflict-free 16-byte row reads
  float* Ks = reinterpret_cast<float*>(xs);             // [Skv][dh]
  float* Vs = Ks + XA_MAX_KV * dh;                      // [Skv][dh]
  float* PPs = Vs + XA_MAX_KV * dh;                     // [Sq][XA_MAX_KV]
  float* dSs = PPs + (size_t)Sq * XA_MAX_KV;            // [Sq][XA_MAX_KV]
  bf16* Qs = reinterpret_cast<bf16*>(dSs + (size_t)Sq * XA_MAX_KV);  // [Sq][dhp]
  bf16* dOs = Qs + (size_t)Sq * dhp;                    // [Sq][dhp]

  const int item = blockIdx.x / H, h = blockIdx.x % H;
  const bf16* kb = k + (long long)item * Skv * ldkv + h * dh;
  const bf16* vb = v + (long long)item * Skv * ldkv + h * dh;
  for (int i = threadIdx.x; i < Skv * dh; i += blockDim.x) {
    const int j = i / dh, d = i % dh;
    Ks[j * dh + d] = __bfloat162float(kb[(long long)j * ldkv + d]);
    Vs[j * dh + d] = __bfloat162float(vb[(long long)j * ldkv + d]);
  }
  const int vec_per_row = dh / 8;
  for (int i = threadIdx.x; i < Sq * vec_per_row; i += blockDim.x) {
    const int r = i / vec_per_row, c = (i % vec_per_row) * 8;
    *reinterpret_cast<uint4*>(Qs + (size_t)r * dhp + c) =
        *reinterpret_cast<const uint4*>(q + ((long long)item * Sq + r) * ldq + h * dh + c);
    *reinterpret_cast<uint4*>(dOs + (size_t)r * dhp + c) =
        *reinterpret_cast<const uint4*>(d_o + ((long long)item * Sq + r) * ldo + h * dh + c);
  }
  __syncthreads();

  for (int r = threadIdx.x; r < Sq; r += blockDim.x) {
    const bf16* qr = Qs + (size_t)r * dhp;
    const bf16* dor = dOs + (size_t)r * dhp;
    float s[XA_MAX_KV], dpp[XA_MAX_KV];
#pragma unroll
    for (int j = 0; j < XA_MAX_KV; ++j) { s[j] = 0.f; dpp[j] = 0.f; }
    for (int d = 0; d < dh; d += 8) {
      float qv[8], gv[8];
      ld8(qr + d, qv);
      ld8(dor + d, gv);
#pragma unroll
      for (int j = 0; j < XA_MAX_KV; ++j) {
        if (j < Skv) {
          const float* kr = Ks + j * dh + d;
          const float* vr = Vs + j * dh + d;
#pragma unroll
          for (int i = 0; i < 8; ++i) { s[j] += qv[i] * kr[i]; dpp[j] += gv[i] * vr[i]; }
        }
      }
    }
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < XA_MAX_KV; ++j)
      if (j < Skv) { s[j] *= pre_scale; mx = fmaxf(mx, s[j]); }
    float den = 0.f;
#pragma unroll
    for (int j = 0; j < XA_MAX_KV; ++j)
      if (j < Skv) { s[j] = __expf(s[j] - mx); den += s[j]; }
    const float invden = 1.f / den;
    float dot = 0.f;  // sum_l dP_l * P_l with dP = post * dPP
#pragma unroll
    for (int j = 0; j < XA_MAX_KV; ++j) {
      s[j] = (j < Skv) ? s[j] * invden : 0.f;  // P
      dot += post_scale * dpp[j] * s[j];
    }
    float ds[XA_MAX_KV];
#pragma unroll
    for (int j = 0; j < XA_MAX_KV; ++j) {
      ds[j] = (j < Skv) ? s[j] * (post_scale * dpp[j] - dot) * pre_scale : 0.f;
      PPs[(size_t)r * XA_MAX_KV + j] = s[j] * post_scale;
      dSs[(size_t)r * XA_MAX_KV + j] = ds[j];
    }
    bf16* dqr = dq + ((long long)item * Sq + r) * lddq + h * dh;
    for (int d = 0; d < dh; d += 8) {
      float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int j = 0; j < XA_MAX_KV; ++j) {
        if (j < Skv) {
          const float* kr = Ks + j * dh + d;
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] += ds[j] * kr[i];
        }
      }
      st8(dqr + d, acc);
    }
  }
  __syncthreads();

  // phase 2: Skv * (dh/8) work items, each reducing over Sq rows for both dK and dV
  for (int wkr = threadIdx.x; wkr < Skv * vec_per_row; wkr += blockDim.x) {
    const int j = wkr / vec_per_row, c = (wkr % vec_per_row) * 8;
    float ak[8] = {0, 0, 0, 0, 0, 0, 0, 0}, av[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int r = 0; r < Sq; ++r) {
      float qv[8], gv[8];
      ld8(Qs + (size_t)r * dhp + c, qv);
      ld8(dOs + (size_t)r * dhp + c, gv);
      const float pp = PPs[(size_t)r * XA_MAX_KV + j], dsv = dSs[(size_t)r * XA_MAX_KV + j];
#pragma unroll
      for (int i = 0; i < 8; ++i) { ak[i] += dsv * qv[i]; av[i] += pp * gv[i]; }
    }
    st8(dk + ((long long)item * Skv + j) * lddkv + h * dh + c, ak);
    st8(dv + ((long long)item * Skv + j) * lddkv + h * dh + c, av);
  }
}

static size_t xattn_bwd_smem(int Sq, int dh) {
  return (size_t)2 * XA_MAX_KV * dh * 4 + (size_t)2 * Sq * XA_MAX_KV * 4 + (size_t)2 * Sq * (dh + 8) * 2;
}

}  // namespace lr2

using namespace lr2;

// tcgen05 kernels (xattn_tc.cu) for the stage shapes (heads of 96, <= 16 keys, 64..256 query rows)
bool lr2_xattn_tc_applicable(int Sq, int Skv, int dh);
int lr2_xattn_tc_fwd(const void* q, long long ldq, const void* k, const void* v, long long ldkv, void* o, long long ldo,
                     int items, int Sq, int Skv, int H, float pre_scale, float post_scale, cudaStream_t stream);
int lr2_xattn_tc_bwd(const void* q, long long ldq, const void* k, const void* v, long long ldkv, const void* d_o,
                     long long ldo, void* dq, long long lddq, void* dk, void* dv, long long lddkv, int items, int Sq,
                     int Skv, int H, float pre_scale, float post_scale, cudaStream_t stream);

static int xattn_check(int items, int Sq, int Skv, int H, int dh, long long ldq, long long ldkv, long long ldo) {
  if (items <= 0 || Sq <= 0 || Skv <= 0 || H <= 0) return LR2_ERR_BAD_SHAPE;
  if (Skv > XA_MAX_KV || dh > XA_MAX_DH || dh % 8) return LR2_ERR_UNSUPPORTED;
  if ((ldq % 8) || (ldkv % 8) || (ldo % 8)) return LR2_ERR_MISALIGNED;
  return LR2_OK;
}

extern "C" int lr2_xattn_fwd(const void* q, long long ldq, const void* k, const void* v, long long ldkv, void* o,
                             long long ldo, int items, int Sq, int Skv, int H, int dh, float pre_scale,
                             float post_scale, void* stream) {
  int rc = xattn_check(items, Sq, Skv, H, dh, ldq, ldkv, ldo);
  if (rc != LR2_OK) return rc;
  if (lr2_xattn_tc_applicable(Sq, Skv, dh))
    return lr2_xattn_tc_fwd(q, ldq, k, v, ldkv, o, ldo, items, Sq, Skv, H, pre_scale, post_scale,
                            reinterpret_cast<cudaStream_t>(stream));
  int threads = ((Sq + 31) / 32) * 32;
  if (threads > 256) threads = 256;
  xattn_fwd_kernel<<<items * H, threads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const bf16*>(q), ldq, reinterpret_cast<const bf16*>(k), reinterpret_cast<const bf16*>(v),
      ldkv, reinterpret_cast<bf16*>(o), ldo, Sq, Skv, H, dh, pre_scale, post_scale); LR2_LAUNCHED(1);
  LR2_RETURN_LAUNCH();
}

extern "C" int lr2_xattn_bwd(const void* q, long long ldq, const void* k, const void* v, long long ldkv,
                             const void* d_o, long long ldo, void* dq, long long lddq, void* dk, void* dv,
                             long long lddkv, int items, int Sq, int Skv, int H, int dh, float pre_scale,
                             float post_scale, void* stream) {
  int rc = xattn_check(items, Sq, Skv, H, dh, ldq, ldkv, ldo);
  if (rc != LR2_OK) return rc;
  if ((lddq % 8) || (lddkv % 8)) return LR2_ERR_MISALIGNED;
  if (lr2_xattn_tc_applicable(Sq, Skv, dh))
    return lr2_xattn_tc_bwd(q, ldq, k, v, ldkv, d_o, ldo, dq, lddq, dk, dv, lddkv, items, Sq, Skv, H, pre_scale,
                            post_scale, reinterpret_cast<cudaStream_t>(stream));
  const size_t smem = xattn_bwd_smem(Sq, dh);
  if (smem > 200 * 1024) return LR2_ERR_UNSUPPORTED;
  static size_t configured = 0;
  if (smem > configured) {
    if (cudaFuncSetAttribute(xattn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return LR2_ERR_CUDA;
    configured = smem;
  }
  int threads = ((Sq + 31) / 32) * 32;
  if (threads > 256) threads = 256;
  if (threads < 64) threads = 64;
  xattn_bwd_kernel<<<items * H, threads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const bf16*>(q), ldq, reinterpret_cast<const bf16*>(k), reinterpret_cast<const bf16*>(v),
      ldkv, reinterpret_cast<const bf16*>(d_o), ldo, reinterpret_cast<bf16*>(dq), lddq, reinterpret_cast<bf16*>(dk),
      reinterpret_cast<bf16*>(dv), lddkv, Sq, Skv, H, dh, pre_scale, post_scale); LR2_LAUNCHED(1);
  LR2_RETURN_LAUNCH();
}
