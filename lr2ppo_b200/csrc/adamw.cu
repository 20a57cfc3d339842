// Multi-tensor AdamW, one launch for every parameter of a model.
// ref: tencentpretrain/utils/optimizers.py:374-402 (HF AdamW, correct_bias=False):
//   m = b1*m + (1-b1)*g ; v = b2*v + (1-b2)*g*g ; p -= lr * m / (sqrt(v) + eps) ; p -= lr*wd*p
// HBM-bound: 28 B/param (fp32 grad) or 26 B/param (bf16 grad) + 2 B/param for the bf16 shadow
// weight the GEMMs read, which is refreshed in the same pass so no separate cast kernel runs.
#include "common.cuh"

namespace lr2 {

constexpr int AW_THREADS = 256;
constexpr int AW_CHUNK = 4096;  // elements per CTA (4 float4 per thread)

__device__ __forceinline__ void adamw_elem(float& p, float g, float& m, float& v, float lr, float b1, float b2,
                                           float eps, float omb1, float omb2, float wd, float lr_decay) {
  m = m * b1 + g * omb1;
  v = v * b2 + (g * g) * omb2;
  const float denom = sqrtf(v) + eps;
  p = p - lr * (m / denom);
  if (wd > 0.f) p = p - (lr_decay * wd) * p;
}

__global__ void __launch_bounds__(AW_THREADS)
adamw_multi_kernel(const void* const* __restrict__ ptrs, const long long* __restrict__ meta,
                   const long long* __restrict__ chunks, const float* __restrict__ hyper) {
  // chunks[2c] = tensor id | (length << 32): length 0 = a default chunk (AW_CHUNK elements or up to the tensor's end);
  // an explicit length (a multiple of 4, <= AW_CHUNK) describes a piece of a COLUMN block of a 2-D parameter
  // (column-sharded out_layer.fc1: each owned row segment is cut into such pieces)
  const long long w0 = chunks[2 * (long long)blockIdx.x];
  const long long t = w0 & 0xFFFFFFFFll;
  const long long len = (w0 >> 32) & 0x7FFFFFFFll;
  const long long off = chunks[2 * (long long)blockIdx.x + 1];
  float* p = (float*)ptrs[6 * t + 0];
  const void* g = ptrs[6 * t + 1];
  float* m = (float*)ptrs[6 * t + 2];
  float* v = (float*)ptrs[6 * t + 3];
  bf16* sh = (bf16*)ptrs[6 * t + 4];
  const long long n = meta[4 * t + 0];
  const float wd = __int_as_float((int)meta[4 * t + 1]);
  const bool g_bf16 = meta[4 * t + 2] != 0;
  const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], omb1 = hyper[4], omb2 = hyper[5],
              gscale = hyper[6], lr_decay = hyper[7];
  const long long end = len ? off + len : min(n, off + AW_CHUNK);
  const bool aligned = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(m) |
                         reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(g)) & 15) == 0 &&
                       (sh == nullptr || (reinterpret_cast<uintptr_t>(sh) & 7) == 0) && (off & 3) == 0;
  auto vec4 = [&](long long i) {
    float4 pv = *reinterpret_cast<const float4*>(p + i);
    float4 mv = *reinterpret_cast<const float4*>(m + i);
    float4 vv = *reinterpret_cast<const float4*>(v + i);
    float4 gv;
    if (g_bf16) {
      const uint2 u = *reinterpret_cast<const uint2*>((const bf16*)g + i);
      const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
      gv = make_float4(a.x, a.y, b.x, b.y);
    } else {
      gv = *reinterpret_cast<const float4*>((const float*)g + i);
    }
    adamw_elem(pv.x, gv.x * gscale, mv.x, vv.x, lr, b1, b2, eps, omb1, omb2, wd, lr_decay);
    adamw_elem(pv.y, gv.y * gscale, mv.y, vv.y, lr, b1, b2, eps, omb1, omb2, wd, lr_decay);
    adamw_elem(pv.z, gv.z * gscale, mv.z, vv.z, lr, b1, b2, eps, omb1, omb2, wd, lr_decay);
    adamw_elem(pv.w, gv.w * gscale, mv.w, vv.w, lr, b1, b2, eps, omb1, omb2, wd, lr_decay);
    *reinterpret_cast<float4*>(p + i) = pv;
    *reinterpret_cast<float4*>(m + i) = mv;
    *reinterpret_cast<float4*>(v + i) = vv;
    if (sh != nullptr) {
      uint2 o;
      o.x = pack_bf16x2(pv.x, pv.y);
      o.y = pack_bf16x2(pv.z, pv.w);
      *reinterpret_cast<uint2*>(sh + i) = o;
    }
  };
  if (aligned && end - off == AW_CHUNK) {
#pragma unroll
    for (int it = 0; it < AW_CHUNK / (AW_THREADS * 4); ++it) vec4(off + (long long)(it * AW_THREADS + threadIdx.x) * 4);
  } else if (aligned && len && (len & 3) == 0) {
    for (long long i = off + (long long)threadIdx.x * 4; i < end; i += AW_THREADS * 4) vec4(i);
  } else {
    for (long long i = off + threadIdx.x; i < end; i += AW_THREADS) {
      float pv = p[i], mv = m[i], vv = v[i];
      const float gv = (g_bf16 ? __bfloat162float(((const bf16*)g)[i]) : ((const float*)g)[i]) * gscale;
      adamw_elem(pv, gv, mv, vv, lr, b1, b2, eps, omb1, omb2, wd, lr_decay);
      p[i] = pv; m[i] = mv; v[i] = vv;
      if (sh != nullptr) sh[i] = __float2bfloat16(pv);
    }
  }
}

}  // namespace lr2

using namespace lr2;

extern "C" int lr2_adamw_chunk_elems(void) { return AW_CHUNK; }

extern "C" int lr2_adamw_multi(const void* const* ptrs, const long long* meta, const long long* chunks,
                               long long num_chunks, const float* hyper, void* stream) {
  if (num_chunks <= 0 || num_chunks > 2147483647LL) return LR2_ERR_BAD_SHAPE;
  if (ptrs == nullptr || meta == nullptr || chunks == nullptr || hyper == nullptr) return LR2_ERR_BAD_SHAPE;
  adamw_multi_kernel<<<(unsigned)num_chunks, AW_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(ptrs, meta,
                                                                                                       chunks, hyper); LR2_LAUNCHED(1);
  LR2_RETURN_LAUNCH();
}
