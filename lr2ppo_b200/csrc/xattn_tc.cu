// Tensor-core XiT cross-attention core (finetune/xit.py:125-148): per (item, head)
//   P = softmax(pre_scale * Q K^T),  O = (post_scale * P) V,   Sq <= 256 query rows, Skv <= 16 keys, heads of 96.
// (softmax FIRST, then the division by sqrt(emb); the reference computes and drops the causal mask.)
// The CUDA-core kernels in xattn.cu are FP32-FMA bound (462 M FMA per forward launch of the stage-3 shape, ~12 us
// floor, 34 us measured; backward 110 us); here the five small GEMMs run on tcgen05 and the kernels are bound by the
// Q / dO / O traffic instead.  Same machinery as mha_tc.cu with the 96-wide head stored as two 128B-swizzled blocks
// of 64 dims (the second half filled):
//   forward : S = Q K^T (UMMA 128 x 16 x 96) -> thread-per-row softmax out of TMEM -> P (bf16, K-major tile)
//             O = P V   (UMMA 128 x 96 x 16, V is the MN-major B operand)   -> O * post/rowsum -> global
//   backward: per 128-query tile: S, dP = dO V^T (N = 16) -> P, dS = P o (post*dP - sum_l post*dP_l P_l) * pre
//             dQ = dS K ; dK += dS^T Q ; dV += (post*P)^T dO  (transposed operands = MN-major views of the P / dS
//             tiles; the 16 keys are rows 0..15 of a UMMA_M = 128 accumulator)
#include <cstdlib>
#include "common.cuh"
#include "tc05.cuh"

namespace lr2 {

constexpr int XT_DH = 96;
constexpr int XT_MAX_KV = 16;
constexpr int XT_MAX_SQ = 256;

__device__ __forceinline__ uint32_t xt_sw(int r, int c) { return (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4); }

// [nrows][96] bf16 rows -> two swizzled blocks of [nrows][64] (block stride `bstride` bytes); rows >= nvalid zero.
__device__ __forceinline__ void xt_load(uint8_t* dst, uint32_t bstride, const bf16* src, long long ld, int nvalid,
                                        int nrows, int tid, int nthreads) {
  const uint32_t d0 = smem_u32(dst);
  for (int idx = tid; idx < nrows * 12; idx += nthreads) {
    const int r = idx / 12, c = idx % 12;                 // 12 chunks of 8 dims
    const bool ok = r < nvalid;
    const bf16* g = src + (ok ? (long long)r * ld + c * 8 : 0);
    const uint32_t off = (uint32_t)(c >> 3) * bstride + xt_sw(r, c & 7);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d0 + off), "l"(g), "r"(ok ? 16 : 0) : "memory");
  }
}
__device__ __forceinline__ void xt_cp_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ------------------------------------------------------------------ forward --
struct XtFwd {     // Q tile 2 x 16 KB | K 2 x 2 KB | V 2 x 2 KB | P 16 KB | barrier
  static constexpr int Q = 0, K = 32768, V = K + 4096, P = V + 4096, BAR = P + 16384, TOTAL = BAR + 64 + 1024;
};

__global__ void __launch_bounds__(128)
xattn_tc_fwd_kernel(const bf16* __restrict__ q, long long ldq, const bf16* __restrict__ k, const bf16* __restrict__ v,
                    long long ldkv, bf16* __restrict__ o, long long ldo, int Sq, int Skv, int H, float pre_scale,
                    float post_scale) {
  extern __shared__ uint8_t xt_raw[];
  uint8_t* sm = xt_raw + ((1024u - (smem_u32(xt_raw) & 1023u)) & 1023u);
  uint8_t* Qs = sm + XtFwd::Q;
  uint8_t* Ks = sm + XtFwd::K;
  uint8_t* Vs = sm + XtFwd::V;
  uint8_t* Ps = sm + XtFwd::P;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + XtFwd::BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int item = blockIdx.x / H, h = blockIdx.x % H;
  const bf16* qb = q + (long long)item * Sq * ldq + h * XT_DH;
  const bf16* kb = k + (long long)item * Skv * ldkv + h * XT_DH;
  const bf16* vb = v + (long long)item * Skv * ldkv + h * XT_DH;
  bf16* ob = o + (long long)item * Sq * ldo + h * XT_DH;
  const int nqt = (Sq + 127) / 128;

  if (warp == 0) {
    if (lane == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  xt_load(Ks, 2048, kb, ldkv, Skv, XT_MAX_KV, tid, 128);
  xt_load(Vs, 2048, vb, ldkv, Skv, XT_MAX_KV, tid, 128);
  xt_load(Qs, 16384, qb, ldq, min(128, Sq), 128, tid, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
  const uint32_t idesc_s = make_idesc_m(128, XT_MAX_KV, false, false);
  const uint32_t idesc_o = make_idesc_m(128, XT_DH, false, true);
  uint32_t phase = 0;
  const float pl2 = pre_scale * 1.4426950408889634f;

  for (int t = 0; t < nqt; ++t) {
    if (t > 0) xt_load(Qs, 16384, qb + (long long)t * 128 * ldq, ldq, min(128, Sq - t * 128), 128, tid, 128);
    xt_cp_wait();
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < XT_DH / 16; ++ks)
        umma_bf16(tmem_base, make_sdesc(smem_u32(Qs) + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024),
                  make_sdesc(smem_u32(Ks) + (ks >> 2) * 2048 + (ks & 3) * 32, 16, 1024), idesc_s, ks > 0 ? 1u : 0u);
      umma_commit(bar);
    }
    mbar_wait(bar, phase); phase ^= 1;
    tc_fence_after();
    const int r = tid, i = t * 128 + r;
    uint32_t a[16];
    tmem_ld16(trow, a);
    tmem_ld_wait();
    float p[16];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      p[j] = __uint_as_float(a[j]) * pl2;
      if (j < Skv) mx = fmaxf(mx, p[j]);
    }
    float l = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) { p[j] = (j < Skv) ? ex2_approx(p[j] - mx) : 0.f; l += p[j]; }
    uint4 u0, u1;
    u0.x = pack_bf16x2(p[0], p[1]); u0.y = pack_bf16x2(p[2], p[3]); u0.z = pack_bf16x2(p[4], p[5]); u0.w = pack_bf16x2(p[6], p[7]);
    u1.x = pack_bf16x2(p[8], p[9]); u1.y = pack_bf16x2(p[10], p[11]); u1.z = pack_bf16x2(p[12], p[13]); u1.w = pack_bf16x2(p[14], p[15]);
    *reinterpret_cast<uint4*>(Ps + xt_sw(r, 0)) = u0;
    *reinterpret_cast<uint4*>(Ps + xt_sw(r, 1)) = u1;
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      umma_bf16(tmem_base + 32, make_sdesc(smem_u32(Ps), 16, 1024), make_sdesc(smem_u32(Vs), 2048, 1024), idesc_o, 0u);
      umma_commit(bar);
    }
    mbar_wait(bar, phase); phase ^= 1;
    tc_fence_after();
    const float inv = post_scale / l;
#pragma unroll
    for (int c0 = 0; c0 < XT_DH; c0 += 32) {
      uint32_t acc[32];
      tmem_ld32(trow + 32u + (uint32_t)c0, acc);
      tmem_ld_wait();
      if (i < Sq) {
        bf16* orow = ob + (long long)i * ldo + c0;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(acc[g * 8 + 0]) * inv, __uint_as_float(acc[g * 8 + 1]) * inv);
          u.y = pack_bf16x2(__uint_as_float(acc[g * 8 + 2]) * inv, __uint_as_float(acc[g * 8 + 3]) * inv);
          u.z = pack_bf16x2(__uint_as_float(acc[g * 8 + 4]) * inv, __uint_as_float(acc[g * 8 + 5]) * inv);
          u.w = pack_bf16x2(__uint_as_float(acc[g * 8 + 6]) * inv, __uint_as_float(acc[g * 8 + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + g * 8) = u;
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
  }
}

// ----------------------------------------------------------------- backward --
struct XtBwd {     // Q tile 2x16 KB | dO tile 2x16 KB | K 4 KB | V 4 KB | PP 16 KB | dS 16 KB | pad 16 KB | barrier
  static constexpr int Q = 0, G = 32768, K = 65536, V = K + 4096, P = V + 4096, DS = P + 16384, PAD = DS + 16384,
                       BAR = PAD + 16384, TOTAL = BAR + 64 + 1024;
};
// TMEM columns: S 0..15, dP 32..47, dQ 64..159, dK 160..255, dV 256..351  -> 512 allocated

__global__ void __launch_bounds__(128)
xattn_tc_bwd_kernel(const bf16* __restrict__ q, long long ldq, const bf16* __restrict__ k, const bf16* __restrict__ v,
                    long long ldkv, const bf16* __restrict__ d_o, long long ldo, bf16* __restrict__ dq, long long lddq,
                    bf16* __restrict__ dk, bf16* __restrict__ dv, long long lddkv, int Sq, int Skv, int H,
                    float pre_scale, float post_scale) {
  extern __shared__ uint8_t xt_raw[];
  uint8_t* sm = xt_raw + ((1024u - (smem_u32(xt_raw) & 1023u)) & 1023u);
  uint8_t* Qs = sm + XtBwd::Q;
  uint8_t* Gs = sm + XtBwd::G;
  uint8_t* Ks = sm + XtBwd::K;
  uint8_t* Vs = sm + XtBwd::V;
  uint8_t* Ps = sm + XtBwd::P;
  uint8_t* dSs = sm + XtBwd::DS;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + XtBwd::BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int item = blockIdx.x / H, h = blockIdx.x % H;
  const bf16* qb = q + (long long)item * Sq * ldq + h * XT_DH;
  const bf16* gb = d_o + (long long)item * Sq * ldo + h * XT_DH;
  const bf16* kb = k + (long long)item * Skv * ldkv + h * XT_DH;
  const bf16* vb = v + (long long)item * Skv * ldkv + h * XT_DH;
  const int nqt = (Sq + 127) / 128;

  if (warp == 0) {
    if (lane == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  xt_load(Ks, 2048, kb, ldkv, Skv, XT_MAX_KV, tid, 128);
  xt_load(Vs, 2048, vb, ldkv, Skv, XT_MAX_KV, tid, 128);
  // the P / dS tiles are read as 128-"key" MN-major operands: keys 16..63 of the block and the block behind it only
  // feed accumulator rows that are discarded, but they must hold finite numbers -> zero once
  for (int idx = tid; idx < (3 * 16384) / 16; idx += 128) reinterpret_cast<uint4*>(Ps)[idx] = make_uint4(0, 0, 0, 0);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
  const uint32_t idesc_s = make_idesc_m(128, XT_MAX_KV, false, false);
  const uint32_t idesc_nn = make_idesc_m(128, XT_DH, false, true);
  const uint32_t idesc_tt = make_idesc_m(128, XT_DH, true, true);
  uint32_t phase = 0;
  const float pl2 = pre_scale * 1.4426950408889634f;

  for (int t = 0; t < nqt; ++t) {
    const int nq = min(128, Sq - t * 128);
    xt_load(Qs, 16384, qb + (long long)t * 128 * ldq, ldq, nq, 128, tid, 128);
    xt_load(Gs, 16384, gb + (long long)t * 128 * ldo, ldo, nq, 128, tid, 128);
    xt_cp_wait();
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < XT_DH / 16; ++ks)
        umma_bf16(tmem_base, make_sdesc(smem_u32(Qs) + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024),
                  make_sdesc(smem_u32(Ks) + (ks >> 2) * 2048 + (ks & 3) * 32, 16, 1024), idesc_s, ks > 0 ? 1u : 0u);
#pragma unroll
      for (int ks = 0; ks < XT_DH / 16; ++ks)
        umma_bf16(tmem_base + 32, make_sdesc(smem_u32(Gs) + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024),
                  make_sdesc(smem_u32(Vs) + (ks >> 2) * 2048 + (ks & 3) * 32, 16, 1024), idesc_s, ks > 0 ? 1u : 0u);
      umma_commit(bar);
    }
    mbar_wait(bar, phase); phase ^= 1;
    tc_fence_after();
    const int r = tid, i = t * 128 + r;
    {
      uint32_t sa[16], da[16];
      tmem_ld16(trow, sa);
      tmem_ld16(trow + 32u, da);
      tmem_ld_wait();
      float p[16];
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        p[j] = __uint_as_float(sa[j]) * pl2;
        if (j < Skv) mx = fmaxf(mx, p[j]);
      }
      float l = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) { p[j] = (j < Skv) ? ex2_approx(p[j] - mx) : 0.f; l += p[j]; }
      const float invl = (i < Sq) ? 1.f / l : 0.f;      // padded query rows contribute nothing
      float dot = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) { p[j] *= invl; dot += post_scale * __uint_as_float(da[j]) * p[j]; }
      float ds[16], pp[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        ds[j] = p[j] * (post_scale * __uint_as_float(da[j]) - dot) * pre_scale;
        pp[j] = p[j] * post_scale;
      }
      uint4 u0, u1, w0, w1;
      u0.x = pack_bf16x2(pp[0], pp[1]); u0.y = pack_bf16x2(pp[2], pp[3]); u0.z = pack_bf16x2(pp[4], pp[5]); u0.w = pack_bf16x2(pp[6], pp[7]);
      u1.x = pack_bf16x2(pp[8], pp[9]); u1.y = pack_bf16x2(pp[10], pp[11]); u1.z = pack_bf16x2(pp[12], pp[13]); u1.w = pack_bf16x2(pp[14], pp[15]);
      w0.x = pack_bf16x2(ds[0], ds[1]); w0.y = pack_bf16x2(ds[2], ds[3]); w0.z = pack_bf16x2(ds[4], ds[5]); w0.w = pack_bf16x2(ds[6], ds[7]);
      w1.x = pack_bf16x2(ds[8], ds[9]); w1.y = pack_bf16x2(ds[10], ds[11]); w1.z = pack_bf16x2(ds[12], ds[13]); w1.w = pack_bf16x2(ds[14], ds[15]);
      *reinterpret_cast<uint4*>(Ps + xt_sw(r, 0)) = u0;
      *reinterpret_cast<uint4*>(Ps + xt_sw(r, 1)) = u1;
      *reinterpret_cast<uint4*>(dSs + xt_sw(r, 0)) = w0;
      *reinterpret_cast<uint4*>(dSs + xt_sw(r, 1)) = w1;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      // dQ_t = dS K  (K = 16 keys: one k-step; B = K MN-major, 96 dims = two 64-wide blocks 2 KB apart)
      umma_bf16(tmem_base + 64, make_sdesc(smem_u32(dSs), 16, 1024), make_sdesc(smem_u32(Ks), 2048, 1024), idesc_nn, 0u);
      // dK += dS^T Q_t, dV += (post P)^T dO_t : A = MN-major view ("128 keys" = this block + the one behind it),
      // K = the 128 query rows of the tile (8 k-steps), B = Q_t / dO_t MN-major (blocks 16 KB apart)
#pragma unroll
      for (int ks = 0; ks < 8; ++ks)
        umma_bf16(tmem_base + 160, make_sdesc(smem_u32(dSs) + ks * 2048, 16384, 1024),
                  make_sdesc(smem_u32(Qs) + ks * 2048, 16384, 1024), idesc_tt, (t > 0 || ks > 0) ? 1u : 0u);
#pragma unroll
      for (int ks = 0; ks < 8; ++ks)
        umma_bf16(tmem_base + 256, make_sdesc(smem_u32(Ps) + ks * 2048, 16384, 1024),
                  make_sdesc(smem_u32(Gs) + ks * 2048, 16384, 1024), idesc_tt, (t > 0 || ks > 0) ? 1u : 0u);
      umma_commit(bar);
    }
    mbar_wait(bar, phase); phase ^= 1;
    tc_fence_after();
#pragma unroll
    for (int c0 = 0; c0 < XT_DH; c0 += 32) {
      uint32_t acc[32];
      tmem_ld32(trow + 64u + (uint32_t)c0, acc);
      tmem_ld_wait();
      if (i < Sq) {
        bf16* drow = dq + ((long long)item * Sq + i) * lddq + h * XT_DH + c0;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(acc[g * 8 + 0]), __uint_as_float(acc[g * 8 + 1]));
          u.y = pack_bf16x2(__uint_as_float(acc[g * 8 + 2]), __uint_as_float(acc[g * 8 + 3]));
          u.z = pack_bf16x2(__uint_as_float(acc[g * 8 + 4]), __uint_as_float(acc[g * 8 + 5]));
          u.w = pack_bf16x2(__uint_as_float(acc[g * 8 + 6]), __uint_as_float(acc[g * 8 + 7]));
          *reinterpret_cast<uint4*>(drow + g * 8) = u;
        }
      }
    }
    tc_fence_before();
  }
  // dK, dV: key j = accumulator row j (warp 0, lanes 0..Skv-1)
  if (warp == 0) {
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
      bf16* dst = (which ? dv : dk) + ((long long)item * Skv + lane) * lddkv + h * XT_DH;
#pragma unroll
      for (int c0 = 0; c0 < XT_DH; c0 += 32) {
        uint32_t acc[32];
        tmem_ld32(trow + (which ? 256u : 160u) + (uint32_t)c0, acc);
        tmem_ld_wait();
        if (lane < Skv) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 u;
            u.x = pack_bf16x2(__uint_as_float(acc[g * 8 + 0]), __uint_as_float(acc[g * 8 + 1]));
            u.y = pack_bf16x2(__uint_as_float(acc[g * 8 + 2]), __uint_as_float(acc[g * 8 + 3]));
            u.z = pack_bf16x2(__uint_as_float(acc[g * 8 + 4]), __uint_as_float(acc[g * 8 + 5]));
            u.w = pack_bf16x2(__uint_as_float(acc[g * 8 + 6]), __uint_as_float(acc[g * 8 + 7]));
            *reinterpret_cast<uint4*>(dst + c0 + g * 8) = u;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace lr2

using namespace lr2;

bool lr2_xattn_tc_applicable(int Sq, int Skv, int dh) {
  static int legacy = -1;
  if (legacy < 0) { const char* e = getenv("LR2_XATTN_LEGACY"); legacy = (e && atoi(e)) ? 1 : 0; }
  return !legacy && dh == XT_DH && Skv <= XT_MAX_KV && Sq >= 64 && Sq <= XT_MAX_SQ;
}

int lr2_xattn_tc_fwd(const void* q, long long ldq, const void* k, const void* v, long long ldkv, void* o, long long ldo,
                     int items, int Sq, int Skv, int H, float pre_scale, float post_scale, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(xattn_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, XtFwd::TOTAL) != cudaSuccess)
      return LR2_ERR_CUDA;
    configured = true;
  }
  xattn_tc_fwd_kernel<<<items * H, 128, XtFwd::TOTAL, stream>>>(
      reinterpret_cast<const bf16*>(q), ldq, reinterpret_cast<const bf16*>(k), reinterpret_cast<const bf16*>(v), ldkv,
      reinterpret_cast<bf16*>(o), ldo, Sq, Skv, H, pre_scale, post_scale);
  LR2_LAUNCHED(1);
  LR2_RETURN_LAUNCH();
}

int lr2_xattn_tc_bwd(const void* q, long long ldq, const void* k, const void* v, long long ldkv, const void* d_o,
                     long long ldo, void* dq, long long lddq, void* dk, void* dv, long long lddkv, int items, int Sq,
                     int Skv, int H, float pre_scale, float post_scale, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(xattn_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, XtBwd::TOTAL) != cudaSuccess)
      return LR2_ERR_CUDA;
    configured = true;
  }
  xattn_tc_bwd_kernel<<<items * H, 128, XtBwd::TOTAL, stream>>>(
      reinterpret_cast<const bf16*>(q), ldq, reinterpret_cast<const bf16*>(k), reinterpret_cast<const bf16*>(v), ldkv,
      reinterpret_cast<const bf16*>(d_o), ldo, reinterpret_cast<bf16*>(dq), lddq, reinterpret_cast<bf16*>(dk),
      reinterpret_cast<bf16*>(dv), lddkv, Sq, Skv, H, pre_scale, post_scale);
  LR2_LAUNCHED(1);
  LR2_RETURN_LAUNCH();
}
