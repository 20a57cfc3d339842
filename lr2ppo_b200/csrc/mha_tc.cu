// Tensor-core multi-headed self-attention core of the TencentPretrain towers (ViT-B/16: S = 197, RoBERTa-base
// S <= 256; heads of 64):   P = softmax(Q K^T * scale + key_bias),  O = dropout(P) V
// ref: tencentpretrain/layers/multi_headed_attn.py:55-76, mask from encoders/transformer_encoder.py:62-68.
//
// tcgen05 / TMEM flash-style kernels, one CTA per (batch, head); S <= 256 means a whole score row fits in TMEM, so the
// softmax is single-pass (no online rescaling) and the S x S matrix never leaves the SM.
//   forward : per 128-query tile   S = Q K^T        (UMMA 128 x NK x 64, accumulators in TMEM columns 0..255)
//             thread-per-row softmax straight out of TMEM -> P~ (bf16, unnormalised, dropout applied) written to
//             shared memory in the canonical K-major 128B-swizzle layout
//                                  O = P~ V          (UMMA 128 x 64 x NK, V is the MN-major B operand)
//             O / rowsum -> global, lse saved for backward.
//   backward: Q, K, V, dO of the head resident in shared memory (one [rows][64] swizzled tile each serves as K-major
//             AND as MN-major operand); for each (128-key half, 128-query tile):
//                 S = Q K^T, dP = dO V^T             (two accumulators, TMEM columns 0..255)
//                 P = exp(S*scale + bias - lse), dS = P o (dP o mask - D) * scale   (8 warps, 64 columns each)
//                 dQ += dS K,  dK += dS^T Q,  dV += (P o mask)^T dO                 (dS^T / P^T are the same
//                 shared-memory tiles read through MN-major descriptors; accumulators in TMEM columns 256..511)
// All five GEMMs of the backward therefore run on the tensor cores without any transposed copy.
// Operand tiles are filled with plain 16-byte loads + manual swizzle (chunk c of row r at c ^ (r & 7)); a
// fence.proxy.async makes them visible to the UMMA reads.  Dropout: one Philox4x32-7 call per 8 keys
// (dropout_mult8), index = ((b*H + h) * 256 + query) * 32 + key/8, regenerated in backward.
#include "common.cuh"
#include "tc05.cuh"

namespace lr2 {

constexpr int TCA_DH = 64;
constexpr int TCA_MAX_S = 256;
constexpr uint32_t TCA_SITE = 0x4D48u;

__device__ __forceinline__ uint32_t sw_off(int r, int c) {          // byte offset of 16-byte chunk c of row r
  return (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4);
}
// [nrows_total][64] bf16 tile, rows >= nvalid zero-filled; src row pitch ld elements.  16-byte cp.async copies
// (zero-fill through src-size 0) keep every chunk of the tile in flight at once; the caller waits with
// cp_async_wait_all() before the proxy fence.
__device__ __forceinline__ void load_tile(uint8_t* dst, const bf16* src, long long ld, int nvalid, int nrows_total,
                                          int tid, int nthreads) {
  const uint32_t d0 = smem_u32(dst);
  for (int idx = tid; idx < nrows_total * 8; idx += nthreads) {
    const int r = idx >> 3, c = idx & 7;
    const bool ok = r < nvalid;
    const bf16* g = src + (ok ? (long long)r * ld + c * 8 : 0);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d0 + sw_off(r, c)), "l"(g), "r"(ok ? 16 : 0)
                 : "memory");
  }
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ uint64_t tca_idx8(long long bh, int i, int j) {
  return ((unsigned long long)bh * TCA_MAX_S + (unsigned long long)i) * (TCA_MAX_S / 8) + (unsigned long long)(j >> 3);
}

// ------------------------------------------------------------------ forward --
// Shared-memory layout of the forward kernel, sized at run time from the padded key count NKp so that short
// sequences keep several CTAs resident per SM (S = 64: 50 KB; S = 197: 103 KB -> two CTAs): Q tile 16 KB | K | V
// (NKp rows of 128 B each) | P~ (at most two 64-key blocks: 128 keys are written and multiplied at a time) |
// key bias | reduction scratch | barrier.
struct FwdLayout {
  int k, v, p, b, red, bar, total;
  __host__ __device__ explicit FwdLayout(int NKp) {
    const int kb = ((NKp * 128 + 1023) / 1024) * 1024;
    const int nblk = (NKp + 63) / 64;
    k = 16384; v = k + kb; p = v + kb; b = p + (nblk < 2 ? nblk : 2) * 16384; red = b + 1024; bar = red + 2048;
    total = bar + 64 + 1024;
  }
};

// 256 threads: warp w works on TMEM lane quadrant w & 3 (query rows) and on column part w >> 2 (one half of the keys
// in the softmax passes, one half of the 64 output dims in the epilogue); the two partial row maxima / row sums meet
// in shared memory.  TMEM: the S accumulator occupies columns [0, NKp); once a 128-key half of S has been turned
// into P~ the O accumulator is built in columns [0, 64) (already consumed), so 64 / 128 / 256 columns suffice.
__global__ void __launch_bounds__(256, 2)
mha_tc_fwd_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k, const bf16* __restrict__ v, long long ld,
                  const float* __restrict__ key_bias, bf16* __restrict__ o, long long ldo, float* __restrict__ lse,
                  int S, int H, float scale, float drop_p, uint32_t thresh16, float dscale,
                  unsigned long long seed_in, const unsigned long long* __restrict__ seed_dev) {
  extern __shared__ uint8_t tca_raw[];
  uint8_t* sm = tca_raw + ((1024u - (smem_u32(tca_raw) & 1023u)) & 1023u);
  const int NKp = (S + 15) & ~15;                      // keys padded to the UMMA N / K granularity
  const FwdLayout L(NKp);
  uint8_t* Qs = sm;
  uint8_t* Ks = sm + L.k;
  uint8_t* Vs = sm + L.v;
  uint8_t* Ps = sm + L.p;
  float* Bs = reinterpret_cast<float*>(sm + L.b);
  float* redm = reinterpret_cast<float*>(sm + L.red);                 // [2][128] partial row maxima
  float* redl = redm + 256;                                            // [2][128] partial row sums
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + L.bar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quad = warp & 3, part = warp >> 2;
  const long long bh = blockIdx.x;
  const int b = (int)(bh / H), h = (int)(bh % H);
  const unsigned long long seed = seed_in + (seed_dev ? *seed_dev : 0ull);
  const long long row0 = (long long)b * S;
  const int nqt = (S + 127) / 128;
  const int nh = (NKp + 127) / 128;                    // 128-key halves
  const uint32_t tcols = NKp <= 64 ? 64u : (NKp <= 128 ? 128u : 256u);
  const int csplit = (((NKp + 31) / 32 + 1) / 2) * 32; // pass 1: part 0 scans [0, csplit), part 1 [csplit, NKp)
  const int cbeg = part ? csplit : 0, cend = part ? NKp : min(csplit, NKp);

  if (warp == 0) {
    if (lane == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(tcols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  load_tile(Ks, k + row0 * ld + h * TCA_DH, ld, S, NKp, tid, 256);
  load_tile(Vs, v + row0 * ld + h * TCA_DH, ld, S, NKp, tid, 256);
  load_tile(Qs, q + row0 * ld + h * TCA_DH, ld, min(128, S), 128, tid, 256);       // first query tile
  for (int j = tid; j < TCA_MAX_S; j += 256) Bs[j] = (key_bias && j < S) ? key_bias[row0 + j] * 1.4426950408889634f : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16);     // this warp's TMEM lane quadrant
  const uint32_t idesc_s = make_idesc_m(128, NKp, false, false);
  const uint32_t idesc_o = make_idesc_m(128, TCA_DH, false, true);
  uint32_t phase = 0;
  const float sl2 = scale * 1.4426950408889634f;
  const int r = quad * 32 + lane;                        // row of the tile; TMEM lane = r

  for (int t = 0; t < nqt; ++t) {
    if (t > 0) load_tile(Qs, q + (row0 + t * 128) * ld + h * TCA_DH, ld, min(128, S - t * 128), 128, tid, 256);
    cp_async_wait_all();
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        umma_bf16(tmem_base, make_sdesc(smem_u32(Qs) + ks * 32, 16, 1024), make_sdesc(smem_u32(Ks) + ks * 32, 16, 1024),
                  idesc_s, ks > 0 ? 1u : 0u);
      umma_commit(bar);
    }
    mbar_wait(bar, phase); phase ^= 1;
    tc_fence_after();

    const int i = t * 128 + r;                           // query index
    // pass 1: partial row maximum of s = acc*scale + bias (log2 domain) over this part's columns
    float mx = -INFINITY;
    for (int c0 = cbeg; c0 < cend; c0 += 32) {
      uint32_t a[32];
      tmem_ld32(trow + (uint32_t)c0, a);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const int j = c0 + e;
        const float sv = fmaf(__uint_as_float(a[e]), sl2, Bs[j & (TCA_MAX_S - 1)]);
        if (j < S) mx = fmaxf(mx, sv);
      }
    }
    redm[part * 128 + r] = mx;
    __syncthreads();
    mx = fmaxf(redm[r], redm[128 + r]);
    // pass 2, one 128-key half at a time: p = 2^(s - mx), partial row sum, dropout, bf16 P~ into the swizzled
    // K-major tile (two blocks of 64 keys), then O (+)= P~ V_half
    float l = 0.f;
    for (int hf = 0; hf < nh; ++hf) {
      const int kbeg = hf * 128, kend = min(NKp, kbeg + 128);
      const int pb = kbeg + part * 64, pe = min(kend, pb + 64);
      for (int c0 = pb; c0 < pe; c0 += 32) {
        uint32_t a[32];
        tmem_ld32(trow + (uint32_t)c0, a);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int j0 = c0 + g * 8;
          if (j0 < pe) {
            float p[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int j = j0 + e;
              const float sv = fmaf(__uint_as_float(a[g * 8 + e]), sl2, Bs[j & (TCA_MAX_S - 1)]);
              p[e] = (j < S) ? ex2_approx(sv - mx) : 0.f;
              l += p[e];
            }
            if (drop_p > 0.f) {
              float m[8];
              dropout_mult8(seed, TCA_SITE, tca_idx8(bh, i & (TCA_MAX_S - 1), j0), thresh16, dscale, m);
#pragma unroll
              for (int e = 0; e < 8; ++e) p[e] *= m[e];
            }
            uint4 u;
            u.x = pack_bf16x2(p[0], p[1]); u.y = pack_bf16x2(p[2], p[3]);
            u.z = pack_bf16x2(p[4], p[5]); u.w = pack_bf16x2(p[6], p[7]);
            const int jl = j0 - kbeg;
            *reinterpret_cast<uint4*>(Ps + (jl >> 6) * 16384 + sw_off(r, (jl & 63) >> 3)) = u;
          }
        }
      }
      fence_proxy_async();
      tc_fence_before();
      __syncthreads();           // every S column of this half has been read: O may overwrite columns [0, 64)
      if (tid == 0) {
        tc_fence_after();
        for (int ks = 0; ks < (kend - kbeg) / 16; ++ks)
          umma_bf16(tmem_base, make_sdesc(smem_u32(Ps) + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024),
                    make_sdesc(smem_u32(Vs) + (kbeg / 16 + ks) * 2048, 8192, 1024), idesc_o,
                    (hf > 0 || ks > 0) ? 1u : 0u);
        umma_commit(bar);
      }
      mbar_wait(bar, phase); phase ^= 1;     // P~ buffer is free again, O holds this half's contribution
      tc_fence_after();
    }
    redl[part * 128 + r] = l;
    __syncthreads();
    l = redl[r] + redl[128 + r];
    const float inv = 1.f / l;
    {
      uint32_t a[32];
      tmem_ld32(trow + (uint32_t)part * 32u, a);
      tmem_ld_wait();
      if (i < S) {
        bf16* orow = o + (row0 + i) * ldo + h * TCA_DH + part * 32;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(a[g * 8 + 0]) * inv, __uint_as_float(a[g * 8 + 1]) * inv);
          u.y = pack_bf16x2(__uint_as_float(a[g * 8 + 2]) * inv, __uint_as_float(a[g * 8 + 3]) * inv);
          u.z = pack_bf16x2(__uint_as_float(a[g * 8 + 4]) * inv, __uint_as_float(a[g * 8 + 5]) * inv);
          u.w = pack_bf16x2(__uint_as_float(a[g * 8 + 6]) * inv, __uint_as_float(a[g * 8 + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + g * 8) = u;
        }
      }
    }
    if (lse != nullptr && part == 0 && i < S) lse[bh * S + i] = (mx + __log2f(l)) * 0.6931471805599453f;   // natural log
    tc_fence_before();      // TMEM reads of this tile are complete before the next tile's MMAs overwrite it
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tcols) : "memory");
  }
}

// ----------------------------------------------------------------- backward --
// Run-time layout of the backward kernel.  Order dS | P | Q | K | V | dO: the MN-major descriptors that read dS^T / P^T
// as 128-"key" operands always cover two 64-key blocks; when only one block exists (S <= 64: "small" mode) the second
// block is whatever follows in shared memory -- finite bf16 data that only feeds discarded accumulator rows.
// small mode (S <= 64, e.g. RoBERTa at 64 tokens): 99 KB of smem and 256 TMEM columns (S/dQ 0..63, dP 64..127,
// dK 128..191, dV 192..255; dQ reuses the consumed S columns) -> two CTAs per SM.
// large mode: 512 columns (S 0..127, dP 128..255, dQ_t 256 + 64 t, dK 384, dV 448), one CTA per SM.
struct BwdLayout {
  int ds, p, q, k, v, g, l, d, b, bar, total;
  bool small;
  uint32_t c_dp, c_dq, c_dk, c_dv, tcols;
  __host__ __device__ explicit BwdLayout(int S) {
    const int nt = (S + 127) / 128;
    small = S <= 64;
    const int pb = small ? 16384 : 32768, tb = nt * 16384;
    ds = 0; p = pb; q = 2 * pb; k = q + tb; v = k + tb; g = v + tb; l = g + tb; d = l + 1024; b = d + 1024;
    bar = b + 1024; total = bar + 64 + 1024;
    c_dp = small ? 64u : 128u; c_dq = small ? 0u : 256u; c_dk = small ? 128u : 384u; c_dv = small ? 192u : 448u;
    tcols = small ? 256u : 512u;
  }
};

__global__ void __launch_bounds__(256, 2)
mha_tc_bwd_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k, const bf16* __restrict__ v, long long ld,
                  const float* __restrict__ key_bias, const bf16* __restrict__ o, const bf16* __restrict__ d_o,
                  long long ldo, const float* __restrict__ lse, bf16* __restrict__ dq, bf16* __restrict__ dk,
                  bf16* __restrict__ dv, long long ldd, int S, int H, float scale, float drop_p, uint32_t thresh16,
                  float dscale, unsigned long long seed_in, const unsigned long long* __restrict__ seed_dev) {
  extern __shared__ uint8_t tca_raw[];
  uint8_t* sm = tca_raw + ((1024u - (smem_u32(tca_raw) & 1023u)) & 1023u);
  const BwdLayout L(S);
  uint8_t* Qs = sm + L.q;
  uint8_t* Ks = sm + L.k;
  uint8_t* Vs = sm + L.v;
  uint8_t* Gs = sm + L.g;
  uint8_t* Ps = sm + L.p;
  uint8_t* dSs = sm + L.ds;
  float* Ls = reinterpret_cast<float*>(sm + L.l);
  float* Ds = reinterpret_cast<float*>(sm + L.d);
  float* Bs = reinterpret_cast<float*>(sm + L.b);
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + L.bar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quad = warp & 3, part = warp >> 2;
  const long long bh = blockIdx.x;
  const int b = (int)(bh / H), h = (int)(bh % H);
  const unsigned long long seed = seed_in + (seed_dev ? *seed_dev : 0ull);
  const long long row0 = (long long)b * S;
  const int nt = (S + 127) / 128;                       // 128-row tiles of queries == of keys

  if (warp == 0) {
    if (lane == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(L.tcols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  load_tile(Qs, q + row0 * ld + h * TCA_DH, ld, S, nt * 128, tid, 256);
  load_tile(Ks, k + row0 * ld + h * TCA_DH, ld, S, nt * 128, tid, 256);
  load_tile(Vs, v + row0 * ld + h * TCA_DH, ld, S, nt * 128, tid, 256);
  load_tile(Gs, d_o + row0 * ldo + h * TCA_DH, ldo, S, nt * 128, tid, 256);
  {
    // D_i = rowsum(dO_i o O_i); thread = row
    const int i = tid;
    float dsum = 0.f;
    if (i < S) {
      const bf16* orow = o + (row0 + i) * ldo + h * TCA_DH;
      const bf16* grow = d_o + (row0 + i) * ldo + h * TCA_DH;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float ov[8], gv[8];
        unpack8(*reinterpret_cast<const uint4*>(orow + c * 8), ov);
        unpack8(*reinterpret_cast<const uint4*>(grow + c * 8), gv);
#pragma unroll
        for (int e = 0; e < 8; ++e) dsum = fmaf(ov[e], gv[e], dsum);
      }
    }
    Ds[i] = dsum;
    Ls[i] = (i < S) ? lse[bh * S + i] * 1.4426950408889634f : 0.f;      // log2 domain
    Bs[i] = (key_bias && i < S) ? key_bias[row0 + i] * 1.4426950408889634f : 0.f;
  }
  cp_async_wait_all();
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16);
  const uint32_t idesc_nn = make_idesc_m(128, TCA_DH, false, true);    // dQ = dS K
  const uint32_t idesc_tt = make_idesc_m(128, TCA_DH, true, true);     // dK = dS^T Q, dV = P^T dO
  uint32_t phase = 0;
  const float sl2 = scale * 1.4426950408889634f;
  const int r = quad * 32 + lane;                       // row of the 128-row tile handled by this thread

  for (int kh = 0; kh < nt; ++kh) {
    const int nk = min(128, S - kh * 128);
    const int NKp = (nk + 15) & ~15;
    const uint32_t idesc_s = make_idesc_m(128, NKp, false, false);
    for (int t = 0; t < nt; ++t) {
      if (tid == 0) {
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_bf16(tmem_base, make_sdesc(smem_u32(Qs) + t * 16384 + ks * 32, 16, 1024),
                    make_sdesc(smem_u32(Ks) + kh * 16384 + ks * 32, 16, 1024), idesc_s, ks > 0 ? 1u : 0u);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_bf16(tmem_base + L.c_dp, make_sdesc(smem_u32(Gs) + t * 16384 + ks * 32, 16, 1024),
                    make_sdesc(smem_u32(Vs) + kh * 16384 + ks * 32, 16, 1024), idesc_s, ks > 0 ? 1u : 0u);
        umma_commit(bar);
      }
      mbar_wait(bar, phase); phase ^= 1;
      tc_fence_after();

      const int i = t * 128 + r;
      const bool row_ok = i < S;
      const float li = Ls[i & (TCA_MAX_S - 1)], di = Ds[i & (TCA_MAX_S - 1)];
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        const int c0 = part * 64 + cc * 16;               // column (key within the half) of this 16-wide step
        if (c0 >= NKp) break;                             // warp-uniform
        uint32_t sa[16], da[16];
        tmem_ld16(trow + (uint32_t)c0, sa);
        tmem_ld16(trow + L.c_dp + (uint32_t)c0, da);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const int jl = c0 + g * 8;                      // key within the half
          const int j0 = kh * 128 + jl;                   // key index
          float m[8];
          if (drop_p > 0.f) dropout_mult8(seed, TCA_SITE, tca_idx8(bh, i & (TCA_MAX_S - 1), j0), thresh16, dscale, m);
          float pd[8], ds[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int j = j0 + e;
            float p = ex2_approx(fmaf(__uint_as_float(sa[g * 8 + e]), sl2, Bs[j & (TCA_MAX_S - 1)]) - li);
            if (!row_ok || j >= S) p = 0.f;
            float dp = __uint_as_float(da[g * 8 + e]);
            if (drop_p > 0.f) { dp *= m[e]; pd[e] = p * m[e]; } else { pd[e] = p; }
            ds[e] = p * (dp - di) * scale;
          }
          uint4 u, w;
          u.x = pack_bf16x2(pd[0], pd[1]); u.y = pack_bf16x2(pd[2], pd[3]);
          u.z = pack_bf16x2(pd[4], pd[5]); u.w = pack_bf16x2(pd[6], pd[7]);
          w.x = pack_bf16x2(ds[0], ds[1]); w.y = pack_bf16x2(ds[2], ds[3]);
          w.z = pack_bf16x2(ds[4], ds[5]); w.w = pack_bf16x2(ds[6], ds[7]);
          const uint32_t off = (uint32_t)(jl >> 6) * 16384u + sw_off(r, (jl & 63) >> 3);
          *reinterpret_cast<uint4*>(Ps + off) = u;
          *reinterpret_cast<uint4*>(dSs + off) = w;
        }
      }
      fence_proxy_async();
      tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
        // dQ_t (+)= dS K_kh : A = dS K-major (K = keys), B = K_kh MN-major
        for (int ks = 0; ks < NKp / 16; ++ks)
          umma_bf16(tmem_base + L.c_dq + t * 64, make_sdesc(smem_u32(dSs) + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024),
                    make_sdesc(smem_u32(Ks) + kh * 16384 + ks * 2048, 8192, 1024), idesc_nn,
                    (kh > 0 || ks > 0) ? 1u : 0u);
        // dK_kh (+)= dS^T Q_t, dV_kh (+)= P^T dO_t : A = MN-major view of the [q][keys] tiles (K = the 128 query rows)
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          umma_bf16(tmem_base + L.c_dk, make_sdesc(smem_u32(dSs) + ks * 2048, 16384, 1024),
                    make_sdesc(smem_u32(Qs) + t * 16384 + ks * 2048, 8192, 1024), idesc_tt,
                    (t > 0 || ks > 0) ? 1u : 0u);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          umma_bf16(tmem_base + L.c_dv, make_sdesc(smem_u32(Ps) + ks * 2048, 16384, 1024),
                    make_sdesc(smem_u32(Gs) + t * 16384 + ks * 2048, 8192, 1024), idesc_tt,
                    (t > 0 || ks > 0) ? 1u : 0u);
        umma_commit(bar);
      }
      mbar_wait(bar, phase); phase ^= 1;
      tc_fence_after();
    }
    // dK_kh (warps 0-3) / dV_kh (warps 4-7): key row j = kh*128 + r
    {
      const int j = kh * 128 + r;
      bf16* dst = (part == 0 ? dk : dv) + (row0 + j) * ldd + h * TCA_DH;
      const uint32_t col = part ? L.c_dv : L.c_dk;
#pragma unroll
      for (int c0 = 0; c0 < TCA_DH; c0 += 32) {
        uint32_t a[32];
        tmem_ld32(trow + col + (uint32_t)c0, a);
        tmem_ld_wait();
        if (j < S) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 u;
            u.x = pack_bf16x2(__uint_as_float(a[g * 8 + 0]), __uint_as_float(a[g * 8 + 1]));
            u.y = pack_bf16x2(__uint_as_float(a[g * 8 + 2]), __uint_as_float(a[g * 8 + 3]));
            u.z = pack_bf16x2(__uint_as_float(a[g * 8 + 4]), __uint_as_float(a[g * 8 + 5]));
            u.w = pack_bf16x2(__uint_as_float(a[g * 8 + 6]), __uint_as_float(a[g * 8 + 7]));
            *reinterpret_cast<uint4*>(dst + c0 + g * 8) = u;
          }
        }
      }
      tc_fence_before();
    }
  }
  // dQ_t: warps split the 64 columns in halves of 32
  for (int t = 0; t < nt; ++t) {
    const int i = t * 128 + r;
    uint32_t a[32];
    tmem_ld32(trow + L.c_dq + (uint32_t)t * 64u + (uint32_t)part * 32u, a);
    tmem_ld_wait();
    if (i < S) {
      bf16* dst = dq + (row0 + i) * ldd + h * TCA_DH + part * 32;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 u;
        u.x = pack_bf16x2(__uint_as_float(a[g * 8 + 0]), __uint_as_float(a[g * 8 + 1]));
        u.y = pack_bf16x2(__uint_as_float(a[g * 8 + 2]), __uint_as_float(a[g * 8 + 3]));
        u.z = pack_bf16x2(__uint_as_float(a[g * 8 + 4]), __uint_as_float(a[g * 8 + 5]));
        u.w = pack_bf16x2(__uint_as_float(a[g * 8 + 6]), __uint_as_float(a[g * 8 + 7]));
        *reinterpret_cast<uint4*>(dst + g * 8) = u;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(L.tcols) : "memory");
  }
}

}  // namespace lr2

using namespace lr2;

// Entry points used by lr2_mha_fwd / lr2_mha_bwd (mha.cu) when dh == 64 and S <= 256.
int lr2_mha_tc_fwd(const void* q, const void* k, const void* v, long long ld, const float* key_bias, void* o,
                   long long ldo, float* lse, int B, int S, int H, float scale, float drop_p, unsigned long long seed,
                   const void* seed_dev, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(mha_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             FwdLayout(TCA_MAX_S).total) != cudaSuccess)
      return LR2_ERR_CUDA;
    configured = true;
  }
  const int smem_fwd = FwdLayout((S + 15) & ~15).total;
  mha_tc_fwd_kernel<<<B * H, 256, smem_fwd, stream>>>(
      reinterpret_cast<const bf16*>(q), reinterpret_cast<const bf16*>(k), reinterpret_cast<const bf16*>(v), ld, key_bias,
      reinterpret_cast<bf16*>(o), ldo, lse, S, H, scale, drop_p, dropout_thresh16(drop_p), dropout_scale16(drop_p), seed,
      reinterpret_cast<const unsigned long long*>(seed_dev));
  LR2_LAUNCHED(1);
  LR2_RETURN_LAUNCH();
}

int lr2_mha_tc_bwd(const void* q, const void* k, const void* v, long long ld, const float* key_bias, const void* o,
                   const void* d_o, long long ldo, const float* lse, void* dq, void* dk, void* dv, long long ldd, int B,
                   int S, int H, float scale, float drop_p, unsigned long long seed, const void* seed_dev,
                   cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(mha_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             BwdLayout(TCA_MAX_S).total) != cudaSuccess)
      return LR2_ERR_CUDA;
    configured = true;
  }
  const int smem_bwd = BwdLayout(S).total;
  mha_tc_bwd_kernel<<<B * H, 256, smem_bwd, stream>>>(
      reinterpret_cast<const bf16*>(q), reinterpret_cast<const bf16*>(k), reinterpret_cast<const bf16*>(v), ld, key_bias,
      reinterpret_cast<const bf16*>(o), reinterpret_cast<const bf16*>(d_o), ldo, lse, reinterpret_cast<bf16*>(dq),
      reinterpret_cast<bf16*>(dk), reinterpret_cast<bf16*>(dv), ldd, S, H, scale, drop_p, dropout_thresh16(drop_p),
      dropout_scale16(drop_p), seed, reinterpret_cast<const unsigned long long*>(seed_dev));
  LR2_LAUNCHED(1);
  LR2_RETURN_LAUNCH();
}
