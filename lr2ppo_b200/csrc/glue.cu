// Memory-bound glue kernels around the dense path: cast+gather, concat row copies,
// bias-gradient column sums, the 768->1 heads, positional embedding add.
// All are pure streaming kernels: 16-byte vector accesses, grid sized in multiples of the SM count.
#include "common.cuh"

namespace lr2 {

constexpr int GRID_CAP = 148 * 8;

__device__ __forceinline__ void ld8f(const bf16* p, float (&v)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 t;
  t = unpack_bf16x2(u.x); v[0] = t.x; v[1] = t.y;
  t = unpack_bf16x2(u.y); v[2] = t.x; v[3] = t.y;
  t = unpack_bf16x2(u.z); v[4] = t.x; v[5] = t.y;
  t = unpack_bf16x2(u.w); v[6] = t.x; v[7] = t.y;
}
__device__ __forceinline__ void st8f(bf16* p, const float (&v)[8]) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
  u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}

// ref: finetune/ppo.py:268-271 — text_emb[batch_index, index], fused with fp32->bf16.
__global__ void cast_gather_kernel(const float* __restrict__ src, const long long* __restrict__ index,
                                   bf16* __restrict__ dst, int bs, int T_src, int T_dst, long long row_elems) {
  const long long vec_per_row = row_elems / 8;
  const long long total = (long long)bs * T_dst * vec_per_row;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long slot = i / vec_per_row, c = (i % vec_per_row) * 8;
    const long long b = slot / T_dst, j = slot % T_dst;
    const long long sj = index ? index[b * T_dst + j] : j;
    const float4* s = reinterpret_cast<const float4*>(src + (b * T_src + sj) * row_elems + c);
    const float4 a = s[0], e = s[1];
    const float v[8] = {a.x, a.y, a.z, a.w, e.x, e.y, e.z, e.w};
    st8f(dst + slot * row_elems + c, v);
  }
}

__global__ void rows_copy_kernel(const bf16* __restrict__ src, long long src_gstride, long long src_off,
                                 bf16* __restrict__ dst, long long dst_gstride, long long dst_off, long long groups,
                                 long long rows_per_group, int D, int accumulate) {
  const int vec = D / 8;
  const long long total = groups * rows_per_group * vec;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / vec;
    const int c = (int)(i % vec) * 8;
    const long long g = row / rows_per_group, r = row % rows_per_group;
    const bf16* s = src + (g * src_gstride + src_off + r) * D + c;
    bf16* d = dst + (g * dst_gstride + dst_off + r) * D + c;
    if (accumulate) {
      float a[8], b[8];
      ld8f(s, a);
      ld8f(d, b);
#pragma unroll
      for (int k = 0; k < 8; ++k) a[k] += b[k];
      st8f(d, a);
    } else {
      *reinterpret_cast<uint4*>(d) = *reinterpret_cast<const uint4*>(s);
    }
  }
}

// Column sums: grid (col tiles of 256, row slabs). Thread owns 8 columns... each warp covers 256 columns,
// the block's warps stride over rows; deterministic two-stage reduction.
constexpr int CS_SLABS = 128;   // row slabs (grid.y): enough CTAs to keep > 44 KB of loads in flight per SM
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const bf16* __restrict__ x, long long ldx, long long rows, int cols,
                      float* __restrict__ partials) {
  __shared__ float red[8][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 256 + lane * 8;
  const long long rows_per = (rows + gridDim.y - 1) / gridDim.y;
  const long long r0 = blockIdx.y * rows_per;
  const long long r1 = min(rows, r0 + rows_per);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (c < cols) {
    long long r = r0 + warp;
    for (; r + 24 < r1; r += 32) {              // four independent 16-byte loads in flight per lane
      const uint4 u0 = *reinterpret_cast<const uint4*>(x + r * ldx + c);
      const uint4 u1 = *reinterpret_cast<const uint4*>(x + (r + 8) * ldx + c);
      const uint4 u2 = *reinterpret_cast<const uint4*>(x + (r + 16) * ldx + c);
      const uint4 u3 = *reinterpret_cast<const uint4*>(x + (r + 24) * ldx + c);
      float v0[8], v1[8], v2[8], v3[8];
      unpack8(u0, v0); unpack8(u1, v1); unpack8(u2, v2); unpack8(u3, v3);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += (v0[k] + v1[k]) + (v2[k] + v3[k]);
    }
    for (; r < r1; r += 8) {
      float v[8];
      ld8f(x + r * ldx + c, v);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += v[k];
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) red[warp][lane * 8 + k] = acc[k];
  __syncthreads();
  const int cc = blockIdx.x * 256 + threadIdx.x;
  if (cc < cols) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    partials[(size_t)blockIdx.y * cols + cc] = s;
  }
}
// out[c] (+)= sum over slabs of partials[slab][c].  Block = 32 columns x 8 slab groups: each thread walks every 8th
// slab (16 independent loads for 128 slabs instead of a 128-long dependent chain per column), the 8 group sums meet
// in shared memory and are added in a fixed order (deterministic).
__global__ void __launch_bounds__(256)
colsum_final_kernel(const float* __restrict__ partials, int slabs, int cols, float* __restrict__ out, int accumulate) {
  __shared__ float red[8][33];
  const int cl = threadIdx.x & 31, sg = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  float s = 0.f;
  if (c < cols) {
    int b = sg;
    for (; b + 24 < slabs; b += 32) {
      const float v0 = partials[(size_t)b * cols + c], v1 = partials[(size_t)(b + 8) * cols + c];
      const float v2 = partials[(size_t)(b + 16) * cols + c], v3 = partials[(size_t)(b + 24) * cols + c];
      s += v0; s += v1; s += v2; s += v3;
    }
    for (; b < slabs; b += 8) s += partials[(size_t)b * cols + c];
  }
  red[sg][cl] = s;
  __syncthreads();
  if (sg == 0 && c < cols) {
    float t = accumulate ? out[c] : 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) t += red[g][cl];
    out[c] = t;
  }
}

// head: warp per row.  ref: finetune/ppo.py:228 / :293-295
__global__ void rowdot_fwd_kernel(const bf16* __restrict__ x, long long row_stride, long long row_off,
                                  const float* __restrict__ w, const float* __restrict__ b, float* __restrict__ out,
                                  int rows, int D) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const bf16* xr = x + (r * row_stride + row_off) * D;
  float acc = 0.f;
  for (int c = lane * 8; c < D; c += 256) {
    float v[8];
    ld8f(xr + c, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc += v[k] * __ldg(w + c + k);
  }
  acc = warp_sum(acc);
  if (lane == 0) out[r] = acc + (b ? b[0] : 0.f);
}

// dx[all rows] = selected ? dout[r]*w : 0 ; dw[c] = sum_r dout[r]*x[sel(r), c]; db = sum dout
__global__ void rowdot_bwd_dx_kernel(const float* __restrict__ w, const float* __restrict__ dout,
                                     bf16* __restrict__ dx, long long row_stride, long long row_off, int rows, int D) {
  const int vec = D / 8;
  const long long total = (long long)rows * row_stride * vec;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / vec;
    const int c = (int)(i % vec) * 8;
    const long long r = row / row_stride;
    float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (row % row_stride == row_off) {
      const float g = dout[r];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = g * __ldg(w + c + k);
    }
    st8f(dx + row * D + c, v);
  }
}
__global__ void rowdot_bwd_dw_kernel(const bf16* __restrict__ x, long long row_stride, long long row_off,
                                     const float* __restrict__ dout, float* __restrict__ dw, float* __restrict__ db,
                                     int rows, int D) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < D) {
    float s = 0.f;
    for (int r = 0; r < rows; ++r) s += dout[r] * __bfloat162float(x[(r * row_stride + row_off) * D + c]);
    dw[c] = s;
  }
  if (c == 0 && db != nullptr) {
    float s = 0.f;
    for (int r = 0; r < rows; ++r) s += dout[r];
    db[0] = s;
  }
}

// ref: finetune/ppo.py:286-289
__global__ void add_pos_fwd_kernel(bf16* __restrict__ x, const float* __restrict__ pos, int bs, int T, int D) {
  const int vec = D / 8;
  const long long total = (long long)bs * T * vec;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / vec;
    const int c = (int)(i % vec) * 8;
    const int t = (int)(row % T);
    float v[8];
    ld8f(x + row * D + c, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] += __ldg(pos + (long long)t * D + c + k);
    st8f(x + row * D + c, v);
  }
}
__global__ void add_pos_bwd_kernel(const bf16* __restrict__ dx, float* __restrict__ dpos, int bs, int T, int D) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T * D) return;
  const int t = i / D, c = i % D;
  float s = 0.f;
  for (int b = 0; b < bs; ++b) s += __bfloat162float(dx[((long long)b * T + t) * D + c]);
  dpos[i] = s;
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long long n) {
  const long long n8 = n / 8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8;
       i += (long long)gridDim.x * blockDim.x) {
    const float4* s = reinterpret_cast<const float4*>(src + i * 8);
    const float4 a = s[0], e = s[1];
    const float v[8] = {a.x, a.y, a.z, a.w, e.x, e.y, e.z, e.w};
    st8f(dst + i * 8, v);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n % 8)) {
    const long long i = n8 * 8 + threadIdx.x;
    dst[i] = __float2bfloat16(src[i]);
  }
}
__global__ void cast_bf16_f32_kernel(const bf16* __restrict__ src, float* __restrict__ dst, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = __bfloat162float(src[i]);
}

static int grid_for(long long work_items, int threads) {
  long long b = (work_items + threads - 1) / threads;
  if (b > GRID_CAP) b = GRID_CAP;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace lr2

using namespace lr2;
#define S_(x) reinterpret_cast<cudaStream_t>(x)

extern "C" int lr2_cast_gather_bf16(const float* src, const long long* index, void* dst, int bs, int T_src,
                                    int T_dst, long long row_elems, void* stream) {
  if (bs <= 0 || T_src <= 0 || T_dst <= 0 || row_elems <= 0 || row_elems % 8) return LR2_ERR_BAD_SHAPE;
  if (index == nullptr && T_src != T_dst) return LR2_ERR_BAD_SHAPE;
  const long long total = (long long)bs * T_dst * (row_elems / 8);
  cast_gather_kernel<<<grid_for(total, 256), 256, 0, S_(stream)>>>(src, index, reinterpret_cast<bf16*>(dst), bs,
                                                                    T_src, T_dst, row_elems); LR2_LAUNCHED(1);
  LR2_RETURN_LAUNCH();
}

extern "C" int lr2_rows_copy_bf16(const void* src, long long src_gstride, long long src_off, void* dst,
                                  long long dst_gstride, long long dst_off, long long groups,
                                  long long rows_per_group, int D, int accumulate, void* stream) {
  if (groups <= 0 || rows_per_group <= 0 || D <= 0 || D % 8) return LR2_ERR_BAD_SHAPE;
  const long long total = groups * rows_per_group * (D / 8);
  rows_copy_kernel<<<grid_for(total, 256), 256, 0, S_(stream)>>>(reinterpret_cast<const bf16*>(src), src_gstride,
                                                                  src_off, reinterpret_cast<bf16*>(dst), dst_gstride,
                                                                  dst_off, groups, rows_per_group, D, accumulate); LR2_LAUNCHED(1);
  LR2_RETURN_LAUNCH();
}

extern "C" long long lr2_colsum_partials_floats(int cols) { return (long long)CS_SLABS * cols; }

extern "C" int lr2_colsum_bf16(const void* x, long long ldx, long long rows, int cols, float* out, float* partials,
                               int accumulate, void* stream) {
  if (rows <= 0 || cols <= 0 || cols % 8 || ldx % 8) return LR2_ERR_BAD_SHAPE;
  if (partials == nullptr) return LR2_ERR_BAD_SHAPE;
  int slabs = (int)((rows + 63) / 64);
  if (slabs > CS_SLABS) slabs = CS_SLABS;
  if (slabs < 1) slabs = 1;
  dim3 grid((cols + 255) / 256, slabs);
  colsum_partial_kernel<<<grid, 256, 0, S_(stream)>>>(reinterpret_cast<const bf16*>(x), ldx, rows, cols, partials); LR2_LAUNCHED(1);
  if (cudaGetLastError() != cudaSuccess) return LR2_ERR_CUDA;
  colsum_final_kernel<<<(cols + 31) / 32, 256, 0, S_(stream)>>>(partials, slabs, cols, out, accumulate); LR2_LAUNCHED(1);
  LR2_RETURN_LAUNCH();
}

extern "C" int lr2_rowdot_fwd(const void* x, long long row_stride, long long row_off, const float* w, const float* b,
                              float* out, int rows, int D, void* stream) {
  if (rows <= 0 || D <= 0 || D % 8 || row_stride <= 0 || row_off < 0 || row_off >= row_stride)
    return LR2_ERR_BAD_SHAPE;
  rowdot_fwd_kernel<<<(rows + 7) / 8, 256, 0, S_(stream)>>>(reinterpret_cast<const bf16*>(x), row_stride, row_off, w,
                                                             b, out, rows, D); LR2_LAUNCHED(1);
  LR2_RETURN_LAUNCH();
}

extern "C" int lr2_rowdot_bwd(const void* x, long long row_stride, long long row_off, const float* w,
                              const float* dout, void* dx, float* dw, float* db, int rows, int D, void* stream) {
  if (rows <= 0 || D <= 0 || D % 8 || row_stride <= 0 || row_off < 0 || row_off >= row_stride)
    return LR2_ERR_BAD_SHAPE;
  if (dx != nullptr) {
    const long long total = (long long)rows * row_stride * (D / 8);
    rowdot_bwd_dx_kernel<<<grid_for(total, 256), 256, 0, S_(stream)>>>(w, dout, reinterpret_cast<bf16*>(dx),
                                                                        row_stride, row_off, rows, D); LR2_LAUNCHED(1);
    if (cudaGetLastError() != cudaSuccess) return LR2_ERR_CUDA;
  }
  if (dw != nullptr) {
    rowdot_bwd_dw_kernel<<<(D + 127) / 128, 128, 0, S_(stream)>>>(reinterpret_cast<const bf16*>(x), row_stride,
                                                                   row_off, dout, dw, db, rows, D); LR2_LAUNCHED(1);
  }
  LR2_RETURN_LAUNCH();
}

extern "C" int lr2_add_pos_fwd(void* x, const float* pos, int bs, int T, int D, void* stream) {
  if (bs <= 0 || T <= 0 || D <= 0 || D % 8) return LR2_ERR_BAD_SHAPE;
  const long long total = (long long)bs * T * (D / 8);
  add_pos_fwd_kernel<<<grid_for(total, 256), 256, 0, S_(stream)>>>(reinterpret_cast<bf16*>(x), pos, bs, T, D); LR2_LAUNCHED(1);
  LR2_RETURN_LAUNCH();
}
extern "C" int lr2_add_pos_bwd(const void* dx, float* dpos, int bs, int T, int D, void* stream) {
  if (bs <= 0 || T <= 0 || D <= 0) return LR2_ERR_BAD_SHAPE;
  add_pos_bwd_kernel<<<(T * D + 255) / 256, 256, 0, S_(stream)>>>(reinterpret_cast<const bf16*>(dx), dpos, bs, T, D); LR2_LAUNCHED(1);
  LR2_RETURN_LAUNCH();
}
extern "C" int lr2_cast_f32_to_bf16(const float* src, void* dst, long long n, void* stream) {
  if (n <= 0) return LR2_ERR_BAD_SHAPE;
  if ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) return LR2_ERR_MISALIGNED;
  cast_f32_bf16_kernel<<<grid_for(n / 8 + 1, 256), 256, 0, S_(stream)>>>(src, reinterpret_cast<bf16*>(dst), n); LR2_LAUNCHED(1);
  LR2_RETURN_LAUNCH();
}
extern "C" int lr2_cast_bf16_to_f32(const void* src, float* dst, long long n, void* stream) {
  if (n <= 0) return LR2_ERR_BAD_SHAPE;
  cast_bf16_f32_kernel<<<grid_for(n, 256), 256, 0, S_(stream)>>>(reinterpret_cast<const bf16*>(src), dst, n); LR2_LAUNCHED(1);
  LR2_RETURN_LAUNCH();
}

// bf16 -> bf16 indexed row gather: dst[b, j, :] = src[b, index[b, j], :]  (pooled-feature reuse for the
// reward model's duplicated items, ref: finetune/ppo.py:318-322 with index = [0, 1, pi(0), pi(1)])
namespace lr2 {
__global__ void gather_bf16_kernel(const bf16* __restrict__ src, const long long* __restrict__ index,
                                   bf16* __restrict__ dst, int bs, int T_src, int T_dst, long long row_elems) {
  const long long vec_per_row = row_elems / 8;
  const long long total = (long long)bs * T_dst * vec_per_row;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long slot = i / vec_per_row, c = (i % vec_per_row) * 8;
    const long long b = slot / T_dst, j = slot % T_dst;
    const long long sj = index[b * T_dst + j];
    *reinterpret_cast<uint4*>(dst + slot * row_elems + c) =
        *reinterpret_cast<const uint4*>(src + (b * T_src + sj) * row_elems + c);
  }
}
}  // namespace lr2
extern "C" int lr2_gather_rows_bf16(const void* src, const long long* index, void* dst, int bs, int T_src, int T_dst,
                                    long long row_elems, void* stream) {
  if (bs <= 0 || T_src <= 0 || T_dst <= 0 || row_elems <= 0 || row_elems % 8 || index == nullptr)
    return LR2_ERR_BAD_SHAPE;
  const long long total = (long long)bs * T_dst * (row_elems / 8);
  lr2::gather_bf16_kernel<<<lr2::grid_for(total, 256), 256, 0, S_(stream)>>>(
      reinterpret_cast<const lr2::bf16*>(src), index, reinterpret_cast<lr2::bf16*>(dst), bs, T_src, T_dst, row_elems);
  LR2_LAUNCHED(1);
  LR2_RETURN_LAUNCH();
}

// counter[0] += inc : advances the device-resident dropout seed once per training forward, so that a
// captured CUDA graph draws fresh masks on every replay.
namespace lr2 {
__global__ void bump_counter_kernel(unsigned long long* c, unsigned long long inc) { c[0] += inc; }
}  // namespace lr2
extern "C" int lr2_bump_counter(void* counter, unsigned long long inc, void* stream) {
  if (counter == nullptr) return LR2_ERR_BAD_SHAPE;
  lr2::bump_counter_kernel<<<1, 1, 0, S_(stream)>>>(reinterpret_cast<unsigned long long*>(counter), inc);
  LR2_LAUNCHED(1);
  LR2_RETURN_LAUNCH();
}

// ---- TencentPretrain embedding kernels -------------------------------------------------------------------------
namespace lr2 {
// out[b,s,:] = word[src[b,s],:] + pos[s,:] + segE[seg[b,s],:]   (ref: embeddings/{word,pos,seg}_embedding.py,
// embedding.py:23-30), fp32 tables -> bf16 row
__global__ void embed_sum_kernel(const long long* __restrict__ src, const long long* __restrict__ seg,
                                 const float* __restrict__ word, const float* __restrict__ pos,
                                 const float* __restrict__ segE, bf16* __restrict__ out, long long rows, int S, int D) {
  const int vec = D / 8;
  const long long total = rows * vec;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / vec;
    const int c = (int)(i % vec) * 8;
    const float* w = word + src[r] * D + c;
    const float* p = pos + (r % S) * D + c;
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = w[k] + p[k];
    if (segE != nullptr) {
      const float* g = segE + seg[r] * D + c;
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] += g[k];
    }
    st8f(out + r * D + c, v);
  }
}
// table[idx[r], :] += d[r, :]  (fp32 atomics; embedding-gradient scatter)
__global__ void embed_scatter_kernel(const long long* __restrict__ idx, const bf16* __restrict__ d,
                                     float* __restrict__ table, long long rows, int D) {
  const long long total = rows * D;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / D;
    const int c = (int)(i % D);
    atomicAdd(table + idx[r] * D + c, __bfloat162float(d[i]));
  }
}
// patches[b*P + py*pw + px, (c, ky, kx)] = img[b, c, py*ps+ky, px*ps+kx]  (Conv2d k=s=ps as a GEMM operand,
// ref: embeddings/patch_embedding.py:18,27)
__global__ void patchify_kernel(const float* __restrict__ img, bf16* __restrict__ out, int B, int C, int Hh, int Ww,
                                int ps) {
  const int ph = Hh / ps, pw = Ww / ps;
  const long long K = (long long)C * ps * ps;
  const long long total = (long long)B * ph * pw * K;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / K;
    const int kk = (int)(i % K);
    const int c = kk / (ps * ps), ky = (kk / ps) % ps, kx = kk % ps;
    const int b = (int)(row / (ph * pw)), p = (int)(row % (ph * pw));
    const int y = (p / pw) * ps + ky, x = (p % pw) * ps + kx;
    out[i] = __float2bfloat16(img[(((long long)b * C + c) * Hh + y) * Ww + x]);
  }
}
// out = x * keep(seed, site, idx) / (1 - p)   (elementwise dropout and its backward; same Philox stream as the
// GEMM epilogues so masks can be regenerated anywhere)
__global__ void dropout_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, long long n, float p,
                               uint32_t thresh, unsigned long long seed_in, unsigned int site,
                               const unsigned long long* __restrict__ seed_dev) {
  const unsigned long long seed = seed_in + (seed_dev ? *seed_dev : 0ull);
  const float sc = dropout_scale16(p);
  const long long n8 = n / 8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8;
       i += (long long)gridDim.x * blockDim.x) {
    float v[8], mult[8];
    ld8f(x + i * 8, v);
    dropout_mult8(seed, site, (unsigned long long)i, thresh, sc, mult);
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] *= mult[k];
    st8f(out + i * 8, v);
  }
}
// out = gelu(bf16(x + bias)), pre = bf16(x + bias): the bias + erf-GELU epilogue of a Linear whose fp32 pre-activation
// sums arrive from elsewhere (K-split out_layer.fc1: the per-rank partial products are summed by a reduce-scatter).
// Same arithmetic as the GEMM epilogue EM_BIAS_GELU: GELU is evaluated on the bf16-rounded pre-activation.
__global__ void bias_gelu_kernel(const float* __restrict__ x, const float* __restrict__ bias, bf16* __restrict__ out,
                                 bf16* __restrict__ pre, long long rows, int D) {
  const long long n8 = rows * (D / 8);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % (D / 8)) * 8;
    const float4 a0 = *reinterpret_cast<const float4*>(x + i * 8), a1 = *reinterpret_cast<const float4*>(x + i * 8 + 4);
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + c)), b1 = __ldg(reinterpret_cast<const float4*>(bias + c + 4));
    float v[8] = {a0.x + b0.x, a0.y + b0.y, a0.z + b0.z, a0.w + b0.w, a1.x + b1.x, a1.y + b1.y, a1.z + b1.z, a1.w + b1.w};
    uint4 pr;
    pr.x = pack_bf16x2(v[0], v[1]); pr.y = pack_bf16x2(v[2], v[3]);
    pr.z = pack_bf16x2(v[4], v[5]); pr.w = pack_bf16x2(v[6], v[7]);
    if (pre != nullptr) *reinterpret_cast<uint4*>(pre + i * 8) = pr;
    float xr[8];
    unpack8(pr, xr);
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = gelu_fast(xr[k]);
    st8f(out + i * 8, v);
  }
}
}  // namespace lr2

extern "C" int lr2_embed_sum(const long long* src, const long long* seg, const float* word, const float* pos,
                             const float* seg_table, void* out_bf16, long long rows, int S, int D, void* stream) {
  if (rows <= 0 || S <= 0 || D <= 0 || D % 8) return LR2_ERR_BAD_SHAPE;
  lr2::embed_sum_kernel<<<lr2::grid_for(rows * (D / 8), 256), 256, 0, S_(stream)>>>(
      src, seg, word, pos, seg_table, reinterpret_cast<lr2::bf16*>(out_bf16), rows, S, D);
  LR2_LAUNCHED(1);
  LR2_RETURN_LAUNCH();
}
extern "C" int lr2_embed_scatter_add(const long long* idx, const void* d_bf16, float* table, long long rows, int D,
                                     void* stream) {
  if (rows <= 0 || D <= 0) return LR2_ERR_BAD_SHAPE;
  lr2::embed_scatter_kernel<<<lr2::grid_for(rows * D, 256), 256, 0, S_(stream)>>>(
      idx, reinterpret_cast<const lr2::bf16*>(d_bf16), table, rows, D);
  LR2_LAUNCHED(1);
  LR2_RETURN_LAUNCH();
}
extern "C" int lr2_patchify(const float* img, void* out_bf16, int B, int C, int Hh, int Ww, int ps, void* stream) {
  if (B <= 0 || C <= 0 || ps <= 0 || Hh % ps || Ww % ps) return LR2_ERR_BAD_SHAPE;
  const long long total = (long long)B * C * Hh * Ww;
  lr2::patchify_kernel<<<lr2::grid_for(total, 256), 256, 0, S_(stream)>>>(img, reinterpret_cast<lr2::bf16*>(out_bf16),
                                                                         B, C, Hh, Ww, ps);
  LR2_LAUNCHED(1);
  LR2_RETURN_LAUNCH();
}
extern "C" int lr2_dropout_bf16(const void* x, void* out, long long n, float p, unsigned long long seed,
                                unsigned int site, const void* seed_dev, void* stream) {
  if (n <= 0 || n % 8 || p < 0.f || p >= 1.f) return LR2_ERR_BAD_SHAPE;
  lr2::dropout_kernel<<<lr2::grid_for(n / 8, 256), 256, 0, S_(stream)>>>(
      reinterpret_cast<const lr2::bf16*>(x), reinterpret_cast<lr2::bf16*>(out), n, p, lr2::dropout_thresh16(p), seed, site,
      reinterpret_cast<const unsigned long long*>(seed_dev));
  LR2_LAUNCHED(1);
  LR2_RETURN_LAUNCH();
}

/* out = gelu(bf16(x + bias)), pre (optional) = bf16(x + bias); x fp32 [rows, D], D % 8 == 0 */
extern "C" int lr2_bias_gelu_rows(const float* x, const float* bias, void* out_bf16, void* pre_bf16, long long rows, int D,
                                  void* stream) {
  if (rows <= 0 || D <= 0 || D % 8) return LR2_ERR_BAD_SHAPE;
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(bias) | reinterpret_cast<uintptr_t>(out_bf16) |
       reinterpret_cast<uintptr_t>(pre_bf16)) & 15)
    return LR2_ERR_MISALIGNED;
  lr2::bias_gelu_kernel<<<lr2::grid_for(rows * (D / 8), 256), 256, 0, S_(stream)>>>(
      x, bias, reinterpret_cast<lr2::bf16*>(out_bf16), reinterpret_cast<lr2::bf16*>(pre_bf16), rows, D);
  LR2_LAUNCHED(1);
  LR2_RETURN_LAUNCH();
}
