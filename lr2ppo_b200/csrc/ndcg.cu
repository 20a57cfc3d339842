// NDCG@k: segmented (per-query) shared-memory bitonic sort + DCG with the reference's exact
// arithmetic.  ref: ndcg.py:28-32,54-65; callers finetune/ppo.py:651-659 (torch.sort descending,
// gold[idx], ideal = sort(gold) descending).
//
// Bit-exactness contract: gain_i = float(int64(2**rel_i - 1)), term_i = gain_i / log2_table[i]
// (IEEE fp32 division), dcg = ((0 + term_0) + term_1) + ... strictly sequential in fp32,
// ndcg_k = ideal_k <= 1e-6f ? 1 : pred_k / ideal_k.  Only the sort is parallel.
// Algorithmic HBM bytes: N*(4+8) in + 4*nk out per query (+8N when `order` is requested).
#include "common.cuh"

namespace lr2 {

__device__ __forceinline__ unsigned long long score_key(float s, unsigned int idx) {
  s = s + 0.0f;  // -0.0 -> +0.0 so that both zeros tie (torch.sort semantics)
  unsigned int u = __float_as_uint(s);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // ascending-order-preserving
  u = ~u;                                           // descending
  return ((unsigned long long)u << 32) | idx;      // ties: lower index first (stable)
}
__device__ __forceinline__ unsigned long long label_key(long long l) {
  const unsigned long long asc = (unsigned long long)l ^ 0x8000000000000000ull;
  return ~asc;  // descending labels
}
__device__ __forceinline__ long long label_from_key(unsigned long long k) {
  return (long long)((~k) ^ 0x8000000000000000ull);
}
__device__ __forceinline__ float gain_of(long long rel) {
  // int64 2**rel - 1 with wrap-around like torch integer pow, then int64 -> fp32 (round to nearest)
  // (negative exponents raise in torch; they and rel >= 64 are defined here as 2**rel == 0)
  long long g;
  if (rel < 0 || rel >= 64) g = -1;
  else g = (long long)((1ull << rel) - 1ull);
  return (float)g;
}

constexpr int NDCG_MAX_K = 32;

template <typename K>
__device__ __forceinline__ void bitonic_sort(K* keys, int npad) {
  for (int k = 2; k <= npad; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < npad; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const K a = keys[i], b = keys[ixj];
          const bool up = ((i & k) == 0);
          if ((a > b) == up) { keys[i] = b; keys[ixj] = a; }
        }
      }
      __syncthreads();
    }
  }
}

__global__ void ndcg_kernel(const float* __restrict__ scores, const long long* __restrict__ labels,
                            const int* __restrict__ lens, int N, long long ld, const long long* __restrict__ ks,
                            int nk, const float* __restrict__ log2_table, float* __restrict__ ndcg,
                            long long* __restrict__ order, int npad) {
  extern __shared__ __align__(16) unsigned char nsm[];
  unsigned long long* skey = reinterpret_cast<unsigned long long*>(nsm);  // [npad] (score,idx) keys
  unsigned long long* lkey = skey + npad;                                  // [npad] label keys
  float* tp = reinterpret_cast<float*>(lkey + npad);                       // [npad] predicted terms
  float* ti = tp + npad;                                                   // [npad] ideal terms (fallback path)
  __shared__ int hist[64];        // label histogram (labels in [0, 62]: every realistic relevance scale)
  __shared__ int out_of_range;
  __shared__ float cut_p[NDCG_MAX_K], cut_i[NDCG_MAX_K];
  const int q = blockIdx.x;
  const int n = lens ? min(lens[q], N) : N;
  const float* sq = scores + (long long)q * ld;
  const long long* lq = labels + (long long)q * ld;
  if (threadIdx.x < 64) hist[threadIdx.x] = 0;
  if (threadIdx.x == 0) out_of_range = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < npad; i += blockDim.x) {
    if (i < n) {
      const long long lab = lq[i];
      skey[i] = score_key(sq[i], (unsigned int)i);
      lkey[i] = label_key(lab);
      if (lab >= 0 && lab <= 62) atomicAdd(&hist[(int)lab], 1);
      else out_of_range = 1;
    } else { skey[i] = ~0ull; lkey[i] = ~0ull; }
  }
  __syncthreads();
  bitonic_sort(skey, npad);
  const bool fallback = out_of_range != 0;   // arbitrary int64 labels: sort them too
  if (fallback) bitonic_sort(lkey, npad);
  // sorted cut positions (min(N, k), ascending) and their original slots: rank counting, one thread per k
  // (no serial insertion loop with dependent global loads), and the label histogram's descending exclusive scan by
  // one warp (two bins per lane) instead of a 63-step serial loop
  __shared__ int cut_pos[NDCG_MAX_K], cut_slot[NDCG_MAX_K];
  __shared__ int cut_raw[NDCG_MAX_K];
  __shared__ int hstart[64];      // first ideal position of label L
  if (threadIdx.x < nk) {
    const long long kv = ks[threadIdx.x];
    const long long c = kv < (long long)n ? kv : (long long)n;
    cut_raw[threadIdx.x] = (int)(c < 0 ? 0 : c);
  }
  if (threadIdx.x >= 32 && threadIdx.x < 64) {
    const int lane = threadIdx.x - 32;
    // bins in descending label order: position d = 62 - L; lane handles d = 2*lane, 2*lane + 1
    const int d0 = 2 * lane, d1 = d0 + 1;
    const int h0 = d0 <= 62 ? hist[62 - d0] : 0, h1 = d1 <= 62 ? hist[62 - d1] : 0;
    int incl = h0 + h1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    const int excl = incl - (h0 + h1);
    if (d0 <= 62) hstart[62 - d0] = excl;
    if (d1 <= 62) hstart[62 - d1] = excl + h0;
  }
  __syncthreads();
  if (threadIdx.x < nk) {
    const int mine = cut_raw[threadIdx.x];
    int rank = 0;
    for (int j = 0; j < nk; ++j) {
      const int other = cut_raw[j];
      rank += (other < mine || (other == mine && j < (int)threadIdx.x)) ? 1 : 0;
    }
    cut_pos[rank] = mine; cut_slot[rank] = threadIdx.x;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const unsigned int idx = (unsigned int)(skey[i] & 0xFFFFFFFFull);
    if (order != nullptr) order[(long long)q * ld + i] = idx;
    const float lg = log2_table[i];
    tp[i] = gain_of(lq[idx]) / lg;
    if (fallback) {
      ti[i] = gain_of(label_from_key(lkey[i])) / lg;
    } else {
      // ideal order = labels descending: position i holds the label L with hstart[L] <= i < hstart[L] + hist[L]
      int lo = 0, hi = 62;                       // hstart is non-increasing in L
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (hstart[mid] <= i) hi = mid; else lo = mid + 1; }
      // lo = smallest L with hstart[L] <= i; skip empty bins above the true one
      int L = lo;
      while (L < 62 && hist[L] == 0) ++L;
      ti[i] = gain_of((long long)L) / lg;
    }
  }
  __syncthreads();
  // strictly sequential fp32 sums (two lists -> two warps), walked cut to cut
  if (threadIdx.x == 0 || threadIdx.x == 32) {
    const float* t = threadIdx.x == 0 ? tp : ti;
    float* outc = threadIdx.x == 0 ? cut_p : cut_i;
    float acc = 0.f;
    int i = 0;
    for (int j = 0; j < nk; ++j) {
      const int c = cut_pos[j];
      // the additions stay strictly left-to-right (bit-exactness); the shared-memory loads are batched so that
      // their latency is paid once per 8 terms instead of once per term
      for (; i + 8 <= c; i += 8) {
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = t[i + e];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc = acc + v[e];
      }
      for (; i < c; ++i) acc = acc + t[i];
      outc[cut_slot[j]] = acc;
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < nk; j += blockDim.x) {
    const float p = cut_p[j], t = cut_i[j];
    ndcg[(long long)q * nk + j] = (t <= 1e-6f) ? 1.0f : p / t;
  }
}

}  // namespace lr2

using namespace lr2;

extern "C" int lr2_ndcg_at_k(const float* scores, const long long* labels, const int* lens, int B, int N, long long ld,
                             const long long* ks, int nk, const float* log2_table, float* ndcg, long long* order,
                             void* stream) {
  if (B <= 0 || N <= 0 || nk <= 0 || ld < N) return LR2_ERR_BAD_SHAPE;
  if (N > 4096 || nk > lr2::NDCG_MAX_K) return LR2_ERR_UNSUPPORTED;
  int npad = 2;
  while (npad < N) npad <<= 1;
  const size_t smem = (size_t)npad * (8 + 8 + 4 + 4);
  static size_t configured = 40 * 1024;   // static shared memory of the kernel counts against the 48 KB default
  if (smem > configured) {
    if (cudaFuncSetAttribute(ndcg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return LR2_ERR_CUDA;
    configured = smem;
  }
  int threads = npad / 2;
  if (threads < 32) threads = 64;
  if (threads > 512) threads = 512;
  if (threads < 64) threads = 64;
  ndcg_kernel<<<B, threads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(scores, labels, lens, N, ld, ks, nk,
                                                                            log2_table, ndcg, order, npad); LR2_LAUNCHED(1);
  LR2_RETURN_LAUNCH();
}

// Pre-sorted variant with the reference meter's own signature (ndcg.py:54-65): both relevance lists are
// given in rank order; one thread per (query, list) runs the sequential fp32 DCG.
namespace lr2 {
__global__ void ndcg_presorted_kernel(const long long* __restrict__ pred, const long long* __restrict__ ideal,
                                      const int* __restrict__ lens, int B, int N, const long long* __restrict__ ks,
                                      int nk, const float* __restrict__ log2_table, float* __restrict__ ndcg,
                                      float* __restrict__ scratch /* [B][2][nk] */) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < 2 * B) {
    const int q = t >> 1, which = t & 1;
    const int n = lens ? min(lens[q], N) : N;
    const long long* rel = (which ? ideal : pred) + (long long)q * N;
    float* out = scratch + ((long long)q * 2 + which) * nk;
    // ks need not be sorted: one sequential pass per k would be O(nk*N); instead walk once and
    // record the running sum whenever i+1 equals a cut.
    for (int j = 0; j < nk; ++j) out[j] = 0.f;
    float acc = 0.f;
    for (int i = 0; i < n; ++i) {
      acc = acc + gain_of(rel[i]) / log2_table[i];
      for (int j = 0; j < nk; ++j) {
        const long long cut = ks[j] < (long long)n ? ks[j] : (long long)n;
        if ((long long)(i + 1) == cut) out[j] = acc;
      }
    }
  }
}
__global__ void ndcg_ratio_kernel(const float* __restrict__ scratch, int B, int nk, float* __restrict__ ndcg) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B * nk) return;
  const int q = t / nk, j = t % nk;
  const float p = scratch[((long long)q * 2 + 0) * nk + j], i = scratch[((long long)q * 2 + 1) * nk + j];
  ndcg[t] = (i <= 1e-6f) ? 1.0f : p / i;
}
}  // namespace lr2

extern "C" int lr2_ndcg_presorted(const long long* pred_rel, const long long* true_rel, const int* lens, int B, int N,
                                  const long long* ks, int nk, const float* log2_table, float* ndcg, float* scratch,
                                  void* stream) {
  if (B <= 0 || N <= 0 || nk <= 0 || scratch == nullptr) return LR2_ERR_BAD_SHAPE;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  lr2::ndcg_presorted_kernel<<<(2 * B + 63) / 64, 64, 0, s>>>(pred_rel, true_rel, lens, B, N, ks, nk, log2_table,
                                                             ndcg, scratch); LR2_LAUNCHED(1);
  if (cudaGetLastError() != cudaSuccess) return LR2_ERR_CUDA;
  lr2::ndcg_ratio_kernel<<<(B * nk + 127) / 128, 128, 0, s>>>(scratch, B, nk, ndcg); LR2_LAUNCHED(1);
  LR2_RETURN_LAUNCH();
}
