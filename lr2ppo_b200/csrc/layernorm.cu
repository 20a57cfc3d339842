// LayerNorm forward / backward, warp per row, bf16 I/O, fp32 math.
// mode 0 = torch nn.LayerNorm (ref: finetune/xit.py:31-41,71-74,96-100),
// mode 1 = TencentPretrain LayerNorm (ref: tencentpretrain/layers/layer_norm.py:16-21).
// HBM-bound: 2*rows*D*2 bytes fwd; the row stays in registers between the
// statistics pass and the normalise pass so x is read exactly once.
#include "common.cuh"

namespace lr2 {

constexpr int LN_MAX_CHUNKS = 4;  // D <= 1024 (8 bf16 per lane per chunk)
constexpr int LN_WARPS = 8;
constexpr int LN_BWD_BLOCKS = 296;

__device__ __forceinline__ long long regroup(long long row, int g_in, int g_out, int g_off) {
  return g_in <= 0 ? row : (row / g_in) * g_out + (row % g_in) + g_off;
}

__device__ __forceinline__ void load8(const bf16* p, float (&v)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 t;
  t = unpack_bf16x2(u.x); v[0] = t.x; v[1] = t.y;
  t = unpack_bf16x2(u.y); v[2] = t.x; v[3] = t.y;
  t = unpack_bf16x2(u.z); v[4] = t.x; v[5] = t.y;
  t = unpack_bf16x2(u.w); v[6] = t.x; v[7] = t.y;
}
__device__ __forceinline__ void store8(bf16* p, const float (&v)[8]) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
  u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}

__global__ void __launch_bounds__(LN_WARPS * 32)
ln_fwd_kernel(const bf16* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
              bf16* __restrict__ y, float* __restrict__ stats, long long rows, int D, float eps, int mode, int g_in,
              int g_out, int g_off) {
  const int lane = threadIdx.x & 31;
  const int nch = D / 256 + ((D % 256) ? 1 : 0);
  for (long long row = (long long)blockIdx.x * LN_WARPS + (threadIdx.x >> 5); row < rows;
       row += (long long)gridDim.x * LN_WARPS) {
    const bf16* xr = x + row * D;
    float v[LN_MAX_CHUNKS][8];
    float sum = 0.f;
#pragma unroll
    for (int ch = 0; ch < LN_MAX_CHUNKS; ++ch) {
      const int c = ch * 256 + lane * 8;
      if (ch < nch && c < D) {
        load8(xr + c, v[ch]);
#pragma unroll
        for (int i = 0; i < 8; ++i) sum += v[ch][i];
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[ch][i] = 0.f;
      }
    }
    sum = warp_sum(sum);
    const float mean = sum / (float)D;
    float sq = 0.f;
#pragma unroll
    for (int ch = 0; ch < LN_MAX_CHUNKS; ++ch) {
      const int c = ch * 256 + lane * 8;
      if (ch < nch && c < D) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { const float d = v[ch][i] - mean; sq += d * d; }
      }
    }
    sq = warp_sum(sq);
    float rinv;
    if (mode == 0) rinv = rsqrtf(sq / (float)D + eps);
    else rinv = 1.0f / (sqrtf(sq / (float)(D - 1)) + eps);
    if (stats != nullptr && lane == 0) { stats[2 * row] = mean; stats[2 * row + 1] = rinv; }
    bf16* yr = y + regroup(row, g_in, g_out, g_off) * D;
#pragma unroll
    for (int ch = 0; ch < LN_MAX_CHUNKS; ++ch) {
      const int c = ch * 256 + lane * 8;
      if (ch < nch && c < D) {
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + c + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + c + 4));
        const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = g[i] * ((v[ch][i] - mean) * rinv) + b[i];
        store8(yr + c, o);
      }
    }
  }
}

// dx = rinv * (g - mean(g) - xhat * sum(g*xhat) * kfac), g = dy*gamma
//   torch:   kfac = 1/D
//   tencent: kfac = 1/((D-1) * (1 - eps*rinv))      (std = 1/rinv - eps)
// Register budget: the row is kept as the RAW 16-byte bf16 vectors (x and dy: 2 x NCH uint4) and xhat / g are
// recomputed from them after the two row reductions, so that with the 2 x NCH x 8 dgamma / dbeta accumulators the
// kernel stays under 128 registers and two 256-thread blocks (16 rows in flight) fit on an SM; the first version
// (fp32 xhat / g arrays sized for D = 1024) needed 166 registers -> 8 warps per SM -> 5x above the HBM floor.
template <int NCH>
__global__ void __launch_bounds__(LN_WARPS * 32, 2)
ln_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const float* __restrict__ gamma,
              const float* __restrict__ stats, const bf16* __restrict__ add, bf16* __restrict__ dx,
              bf16* __restrict__ dxm, float* __restrict__ partials, long long rows, int D, float eps, int mode,
              int g_in, int g_out, int g_off, float drop_p, unsigned long long seed_in, unsigned int site,
              const unsigned long long* __restrict__ seed_dev) {
  const unsigned long long seed = seed_in + (seed_dev ? *seed_dev : 0ull);
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nch = D / 256 + ((D % 256) ? 1 : 0);
  float dg[NCH][8], db[NCH][8];
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
    for (int i = 0; i < 8; ++i) { dg[ch][i] = 0.f; db[ch][i] = 0.f; }
  const uint32_t th = dropout_thresh16(drop_p);
  const float dscale = dropout_scale16(drop_p);

  for (long long row = (long long)blockIdx.x * LN_WARPS + warp; row < rows; row += (long long)gridDim.x * LN_WARPS) {
    const bf16* xr = x + row * D;
    const bf16* dyr = dy + regroup(row, g_in, g_out, g_off) * D;
    const float mean = stats[2 * row], rinv = stats[2 * row + 1];
    uint4 xraw[NCH], draw[NCH];
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {           // all loads of the row first
      const int c = ch * 256 + lane * 8;
      if (ch < nch && c < D) {
        xraw[ch] = *reinterpret_cast<const uint4*>(xr + c);
        draw[ch] = *reinterpret_cast<const uint4*>(dyr + c);
      } else {
        xraw[ch] = make_uint4(0, 0, 0, 0); draw[ch] = make_uint4(0, 0, 0, 0);
      }
    }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      const int c = ch * 256 + lane * 8;
      if (ch < nch && c < D) {
        float xv[8], dv[8];
        unpack8(xraw[ch], xv);
        unpack8(draw[ch], dv);
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + c + 4));
        const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float xh = (xv[i] - mean) * rinv;
          const float g = dv[i] * gm[i];
          s1 += g;
          s2 += g * xh;
          dg[ch][i] += dv[i] * xh;
          db[ch][i] += dv[i];
        }
      }
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    const float mg = s1 / (float)D;
    const float kfac = (mode == 0) ? 1.f / (float)D : 1.f / ((float)(D - 1) * (1.f - eps * rinv));
    const float s2k = s2 * kfac;
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      const int c = ch * 256 + lane * 8;
      if (ch < nch && c < D) {
        float xv[8], dv[8];
        unpack8(xraw[ch], xv);
        unpack8(draw[ch], dv);
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + c + 4));
        const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float xh = (xv[i] - mean) * rinv;          // same expressions as in the first pass
          const float g = dv[i] * gm[i];
          o[i] = rinv * (g - mg - xh * s2k);
        }
        if (add != nullptr) {
          float a[8];
          load8(add + row * D + c, a);
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] += a[i];
        }
        store8(dx + row * D + c, o);
        if (dxm != nullptr) {
          float mult[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
          if (drop_p > 0.f) dropout_mult8(seed, site, (uint64_t)(row * D + c) >> 3, th, dscale, mult);
          // mask the bf16-rounded dx so dxm == mask * dx exactly
          float m[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) m[i] = __bfloat162float(__float2bfloat16(o[i])) * mult[i];
          store8(dxm + row * D + c, m);
        }
      }
    }
  }

  // block-level reduction of dgamma / dbeta partials (deterministic order)
  extern __shared__ float sred[];  // [LN_WARPS][2][D]
  float* mine = sred + (size_t)warp * 2 * D;
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    const int c = ch * 256 + lane * 8;
    if (ch < nch && c < D) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { mine[c + i] = dg[ch][i]; mine[D + c + i] = db[ch][i]; }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * D; c += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < LN_WARPS; ++w) s += sred[(size_t)w * 2 * D + c];
    partials[(size_t)blockIdx.x * 2 * D + c] = s;
  }
}

// dgamma / dbeta = sum over the per-block partial rows.  Block = 32 columns x 8 row groups (every thread walks every
// 8th partial row: 37 independent loads for 296 rows instead of one 296-long dependent chain), fixed-order combine.
__global__ void __launch_bounds__(256)
ln_bwd_reduce_kernel(const float* __restrict__ partials, int nblocks, int D, float* __restrict__ dgamma,
                     float* __restrict__ dbeta) {
  __shared__ float red[8][33];
  const int cl = threadIdx.x & 31, sg = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  float s = 0.f;
  if (c < 2 * D) {
    int b = sg;
    for (; b + 24 < nblocks; b += 32) {
      const float v0 = partials[(size_t)b * 2 * D + c], v1 = partials[(size_t)(b + 8) * 2 * D + c];
      const float v2 = partials[(size_t)(b + 16) * 2 * D + c], v3 = partials[(size_t)(b + 24) * 2 * D + c];
      s += v0; s += v1; s += v2; s += v3;
    }
    for (; b < nblocks; b += 8) s += partials[(size_t)b * 2 * D + c];
  }
  red[sg][cl] = s;
  __syncthreads();
  if (sg == 0 && c < 2 * D) {
    float t = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) t += red[g][cl];
    if (c < D) dgamma[c] = t;
    else dbeta[c - D] = t;
  }
}

}  // namespace lr2

using namespace lr2;

static int ln_blocks(long long rows, int cap) {
  long long b = (rows + LN_WARPS - 1) / LN_WARPS;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

extern "C" int lr2_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* stats,
                                 long long rows, int D, float eps, int mode, int g_in, int g_out, int g_off,
                                 void* stream) {
  if (rows <= 0 || D <= 0 || D % 8 || D > 256 * LN_MAX_CHUNKS) return LR2_ERR_BAD_SHAPE;
  if (mode != 0 && mode != 1) return LR2_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) return LR2_ERR_MISALIGNED;
  ln_fwd_kernel<<<ln_blocks(rows, 148 * 16), LN_WARPS * 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const bf16*>(x), gamma, beta, reinterpret_cast<bf16*>(y), stats, rows, D, eps, mode, g_in,
      g_out, g_off); LR2_LAUNCHED(1);
  LR2_RETURN_LAUNCH();
}

extern "C" long long lr2_layernorm_bwd_partials_floats(int D) { return (long long)LN_BWD_BLOCKS * 2 * D; }

extern "C" int lr2_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* stats,
                                 const void* add, void* dx, void* dxm, float* dgamma, float* dbeta, float* partials,
                                 long long rows, int D, float eps, int mode, int g_in, int g_out, int g_off,
                                 float drop_p, unsigned long long seed, unsigned int site, const void* seed_dev,
                                 void* stream) {
  if (rows <= 0 || D <= 0 || D % 8 || D > 256 * LN_MAX_CHUNKS) return LR2_ERR_BAD_SHAPE;
  if (mode != 0 && mode != 1) return LR2_ERR_UNSUPPORTED;
  if (partials == nullptr || stats == nullptr) return LR2_ERR_BAD_SHAPE;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int nb = ln_blocks(rows, LN_BWD_BLOCKS);
  const size_t smem = (size_t)LN_WARPS * 2 * D * sizeof(float);
  const int nchunks = (D + 255) / 256;
#define LR2_LN_BWD(NCH_)                                                                                             \
  do {                                                                                                               \
    static bool configured = false;                                                                                  \
    if (!configured) {                                                                                               \
      if (cudaFuncSetAttribute(ln_bwd_kernel<NCH_>, cudaFuncAttributeMaxDynamicSharedMemorySize,                     \
                               LN_WARPS * 2 * 1024 * 4) != cudaSuccess)                                             \
        return LR2_ERR_CUDA;                                                                                         \
      configured = true;                                                                                             \
    }                                                                                                                \
    ln_bwd_kernel<NCH_><<<nb, LN_WARPS * 32, smem, s>>>(                                                            \
        reinterpret_cast<const bf16*>(dy), reinterpret_cast<const bf16*>(x), gamma, stats,                           \
        reinterpret_cast<const bf16*>(add), reinterpret_cast<bf16*>(dx), reinterpret_cast<bf16*>(dxm), partials,    \
        rows, D, eps, mode, g_in, g_out, g_off, drop_p, seed, site,                                                  \
        reinterpret_cast<const unsigned long long*>(seed_dev));                                                      \
  } while (0)
  if (nchunks <= 1) LR2_LN_BWD(1);
  else if (nchunks == 2) LR2_LN_BWD(2);
  else if (nchunks == 3) LR2_LN_BWD(3);
  else LR2_LN_BWD(4);
#undef LR2_LN_BWD
  LR2_LAUNCHED(1);
  if (cudaGetLastError() != cudaSuccess) return LR2_ERR_CUDA;
  ln_bwd_reduce_kernel<<<(2 * D + 31) / 32, 256, 0, s>>>(partials, nb, D, dgamma, dbeta); LR2_LAUNCHED(1);
  LR2_RETURN_LAUNCH();
}
