// Linear-pass fused weight gradient + AdamW for a Linear weight with a SHORT reduction dimension
// (out_layer.fc1: [3072, 162816] fp32, gradient = dY[rows, out]^T X[rows, in] with rows = items = 48).
//
// Why a second implementation next to the tcgen05 one (gemm_sm100.cu, lr2_gemm_wgrad_adamw): with K = 48 the
// gradient costs 48 FMA per element while the optimizer moves 26 bytes per element, so the pass is an AdamW pass
// (HBM-bound) that happens to need a tiny GEMM, not a GEMM with an optimizer epilogue.  Both operands are small
// (X 15.6 MB, dY 0.3 MB) and stay L2-resident, so every warp recomputes its own 16 x 128 gradient tile with
// mma.sync.m16n8k16 (bf16 in, fp32 accumulate; the fragments are loaded straight from global memory, no shared memory,
// no barriers) and feeds the accumulator fragments to the update.  p / m / v are touched in 32-byte pieces (4 lanes x
// float2) that tile each row contiguously: 16 rows x 512 B per warp, all loads of a 32-column group issued before the
// first use (24 x 8 B in flight per lane; profiles/r01_store_pattern_inflight.txt: this access shape reaches
// 5.3-5.9 TB/s once >= 4 independent accesses per lane are in flight).
//
// Status: compiled and index-checked on the host (tests/test_wgrad_mma_model_cpu.py); NOT yet run on a GPU.  It is
// reached only with LR2_WGRAD_ADAMW_IMPL=mma; the default stays the tcgen05 implementation.
// ref: finetune/ppo.py:579-580 (loss.backward(); optimizer.step()) restricted to out_layer.fc1.weight;
//      tencentpretrain/utils/optimizers.py:374-402.
#include "common.cuh"

namespace lr2 {

constexpr int WG_THREADS = 256;
constexpr int WG_WARPS = WG_THREADS / 32;
constexpr int WG_WCOLS = 128;                    // columns per warp tile (16 n8 tiles)
constexpr int WG_TCOLS = WG_WARPS * WG_WCOLS;    // columns per CTA tile
constexpr int WG_ROWS = 16;                      // rows per CTA tile (one m16 tile)
constexpr int WG_GROUP = 4;                      // n8 tiles whose p / m / v loads are in flight together

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// {src[k][col], src[k+1][col]} packed low / high (the k-pair one mma fragment register holds); rows >= K read as 0
__device__ __forceinline__ uint32_t ld_kpair(const bf16* __restrict__ src, long long ld, int k, int K, long long col) {
  const unsigned short* s = reinterpret_cast<const unsigned short*>(src);
  const uint32_t lo = k < K ? (uint32_t)__ldg(s + (long long)k * ld + col) : 0u;
  const uint32_t hi = k + 1 < K ? (uint32_t)__ldg(s + (long long)(k + 1) * ld + col) : 0u;
  return lo | (hi << 16);
}

template <int KS>   // k-steps of 16: K <= 16 * KS
__global__ void __launch_bounds__(WG_THREADS, 2)
adamw_wgrad_mma_kernel(const bf16* __restrict__ dY, long long lddy, const bf16* __restrict__ X, long long ldx, int K,
                       int in_f, float* __restrict__ P, float* __restrict__ M, float* __restrict__ V,
                       bf16* __restrict__ S, const float* __restrict__ hyper, float wd, int col_tiles) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;                       // mma fragment coordinates
  const int rt = blockIdx.x / col_tiles, ct = blockIdx.x - rt * col_tiles;   // consecutive CTAs walk along a row panel
  const int r0 = rt * WG_ROWS;
  const long long cw = (long long)ct * WG_TCOLS + warp * WG_WCOLS;
  if (cw >= in_f) return;                                      // in_f % 128 == 0: a warp tile is all in or all out
  const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], omb1 = hyper[4], omb2 = hyper[5],
              gs = hyper[6], lrd = hyper[7];
  const float decay = lrd * wd;

  // A = dY^T tile: A[m][k] = dY[k][r0 + m].  reg0 (m = g, k = 2t..), reg1 (m = g + 8), reg2 / reg3 the same at k + 8.
  uint32_t a[KS][4];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    const int k0 = ks * 16 + 2 * t;
    a[ks][0] = ld_kpair(dY, lddy, k0, K, r0 + g);
    a[ks][1] = ld_kpair(dY, lddy, k0, K, r0 + g + 8);
    a[ks][2] = ld_kpair(dY, lddy, k0 + 8, K, r0 + g);
    a[ks][3] = ld_kpair(dY, lddy, k0 + 8, K, r0 + g + 8);
  }

#pragma unroll 1
  for (int grp = 0; grp < WG_WCOLS / (8 * WG_GROUP); ++grp) {
    const long long c0 = cw + grp * (8 * WG_GROUP);
    // element offsets of this lane's accumulator fragments: rows r0 + g (+8), columns c0 + 8 j + 2 t (+1)
    const long long row_lo = (long long)(r0 + g) * in_f + c0 + 2 * t;
    const long long row_hi = row_lo + 8LL * in_f;
    // 1) every p / m / v load of the group (24 x 8 B per lane) before anything is consumed
    float2 pv[WG_GROUP][2], mv[WG_GROUP][2], vv[WG_GROUP][2];
#pragma unroll
    for (int j = 0; j < WG_GROUP; ++j) {
      pv[j][0] = *reinterpret_cast<const float2*>(P + row_lo + 8 * j);
      pv[j][1] = *reinterpret_cast<const float2*>(P + row_hi + 8 * j);
      mv[j][0] = *reinterpret_cast<const float2*>(M + row_lo + 8 * j);
      mv[j][1] = *reinterpret_cast<const float2*>(M + row_hi + 8 * j);
      vv[j][0] = *reinterpret_cast<const float2*>(V + row_lo + 8 * j);
      vv[j][1] = *reinterpret_cast<const float2*>(V + row_hi + 8 * j);
    }
    // 2) gradient tiles: B[k][n] = X[k][c0 + 8 j + n], reg0 (k = 2t.., n = g), reg1 the same at k + 8
    float acc[WG_GROUP][4];
#pragma unroll
    for (int j = 0; j < WG_GROUP; ++j) {
      acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const int k0 = ks * 16 + 2 * t;
        const uint32_t bb0 = ld_kpair(X, ldx, k0, K, c0 + 8 * j + g);
        const uint32_t bb1 = ld_kpair(X, ldx, k0 + 8, K, c0 + 8 * j + g);
        mma_bf16_16816(acc[j], a[ks], bb0, bb1);
      }
    }
    // 3) AdamW on the fragments (acc[j][0..1]: row g, acc[j][2..3]: row g + 8) and the stores
#pragma unroll
    for (int j = 0; j < WG_GROUP; ++j) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const long long off = (h ? row_hi : row_lo) + 8 * j;
        float pe[2] = {pv[j][h].x, pv[j][h].y}, me[2] = {mv[j][h].x, mv[j][h].y}, ve[2] = {vv[j][h].x, vv[j][h].y};
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const float gr = acc[j][2 * h + i] * gs;
          me[i] = me[i] * b1 + gr * omb1;
          ve[i] = ve[i] * b2 + (gr * gr) * omb2;
          const float denom = sqrtf(ve[i]) + eps;
          float pn = pe[i] - lr * (me[i] / denom);
          if (wd > 0.f) pn = pn - decay * pn;
          pe[i] = pn;
        }
        *reinterpret_cast<float2*>(P + off) = make_float2(pe[0], pe[1]);
        *reinterpret_cast<float2*>(M + off) = make_float2(me[0], me[1]);
        *reinterpret_cast<float2*>(V + off) = make_float2(ve[0], ve[1]);
        if (S != nullptr) *reinterpret_cast<uint32_t*>(S + off) = pack_bf16x2(pe[0], pe[1]);
      }
    }
  }
}

}  // namespace lr2

using namespace lr2;

// Called by lr2_gemm_wgrad_adamw (gemm_sm100.cu) when LR2_WGRAD_ADAMW_IMPL=mma and the shape qualifies.
// Returns LR2_ERR_UNSUPPORTED when it does not (the caller then uses the tcgen05 implementation).
int lr2_adamw_wgrad_mma(const void* dY, long long lddy, const void* X, long long ldx, int rows, int out_f, int in_f,
                        float* param, float* exp_avg, float* exp_avg_sq, void* shadow_bf16, const float* hyper,
                        float weight_decay, cudaStream_t stream) {
  if (rows <= 0 || rows > 64 || (out_f % WG_ROWS) || (in_f % WG_WCOLS)) return LR2_ERR_UNSUPPORTED;
  const int col_tiles = (in_f + WG_TCOLS - 1) / WG_TCOLS;
  const long long blocks = (long long)(out_f / WG_ROWS) * col_tiles;
  if (blocks > 2147483647LL) return LR2_ERR_UNSUPPORTED;
  const bf16* dy = reinterpret_cast<const bf16*>(dY);
  const bf16* x = reinterpret_cast<const bf16*>(X);
  bf16* sh = reinterpret_cast<bf16*>(shadow_bf16);
#define LR2_WG_LAUNCH(KS_)                                                                                          \
  adamw_wgrad_mma_kernel<KS_><<<(unsigned)blocks, WG_THREADS, 0, stream>>>(dy, lddy, x, ldx, rows, in_f, param,     \
                                                                            exp_avg, exp_avg_sq, sh, hyper,        \
                                                                            weight_decay, col_tiles)
  const int ks = (rows + 15) / 16;
  if (ks == 1) LR2_WG_LAUNCH(1);
  else if (ks == 2) LR2_WG_LAUNCH(2);
  else if (ks == 3) LR2_WG_LAUNCH(3);
  else LR2_WG_LAUNCH(4);
#undef LR2_WG_LAUNCH
  LR2_LAUNCHED(1);
  LR2_RETURN_LAUNCH();
}
