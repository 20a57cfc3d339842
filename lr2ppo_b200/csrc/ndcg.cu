// NDCG@k: segmented (per-query) shared-memory bitonic sort + DCG with the reference's exact
// arithmetic.  ref: ndcg.py:28-32,54-65; callers finetune/ppo.py:651-659 (torch.sort descending,
// gold[idx], ideal = sort(gold) descending).
//
// Bit-exactness contract: gain_i = float(int64(2**rel_i - 1)), term_i = gain_i / log2_table[i]
// (IEEE fp32 division), dcg = ((0 + term_0) + term_1) + ... strictly sequential in fp32,
// ndcg_k = ideal_k <= 1e-6f ? 1 : pred_k / ideal_k.  Only the sort is parallel.
// Algorithmic HBM bytes: N*(4+8) in + 4*nk out per query (+8N when `order` is requested).
#include "common.cuh"
#include <stdlib.h>

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

// pair_index != 0 (LR2_NDCG_BLOCK_PAIRS=1, opt-in until it has run on a GPU): thread t takes the t-th compare-exchange
// pair directly, i = ((t & ~(j-1)) << 1) | (t & (j-1)), so no thread idles; otherwise every thread visits an element
// and the upper partner of each pair skips (half of the threads idle in every stage).
template <typename K>
__device__ __forceinline__ void bitonic_sort(K* keys, int npad, int pair_index) {
  for (int k = 2; k <= npad; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (pair_index) {
        for (int t = threadIdx.x; t < (npad >> 1); t += blockDim.x) {
          const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
          const int ixj = i | j;
          const K a = keys[i], b = keys[ixj];
          const bool up = ((i & k) == 0);
          if ((a > b) == up) { keys[i] = b; keys[ixj] = a; }
        }
      } else {
        for (int i = threadIdx.x; i < npad; i += blockDim.x) {
          const int ixj = i ^ j;
          if (ixj > i) {
            const K a = keys[i], b = keys[ixj];
            const bool up = ((i & k) == 0);
            if ((a > b) == up) { keys[i] = b; keys[ixj] = a; }
          }
        }
      }
      __syncthreads();
    }
  }
}

__global__ void ndcg_kernel(const float* __restrict__ scores, const long long* __restrict__ labels,
                            const int* __restrict__ lens, int N, long long ld, const long long* __restrict__ ks,
                            int nk, const float* __restrict__ log2_table, float* __restrict__ ndcg,
                            long long* __restrict__ order, int npad, int pair_index) {
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
  bitonic_sort(skey, npad, pair_index);
  const bool fallback = out_of_range != 0;   // arbitrary int64 labels: sort them too
  if (fallback) bitonic_sort(lkey, npad, pair_index);
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


// ---------------------------------------------------------------------------------------------------------------
// Warp-per-query path (N <= 1024): the whole query lives in one warp's registers, E = npad / 32 keys per lane.
// The sort is the mirror ("flip") form of the bitonic network, in which every compare-exchange leaves the smaller
// key at the lower position, so no direction flags exist: for block size kk = 2, 4, .., npad the first stage pairs
// position p with p ^ (kk - 1) and the following stages pair p with p ^ j for j = kk/4 .. 1.  Positions are blocked
// (p = lane * E + e): strides below E are register-to-register (static indices, fully unrolled), strides of E and
// more are __shfl_xor exchanges with lane ^ (j / E) (flip: lane ^ (kk / E - 1), register E - 1 - e).  No
// __syncthreads anywhere; 40 of the 55 stages of a 1024-key sort never leave the register file.
//
// Labels in [0, 62] (every realistic relevance scale) ride in the key's low byte, so the predicted gains need no
// gather, and the ideal ordering comes from a label histogram (per-lane private byte counters in shared memory:
// no atomics, reduced with dp4a).  Any other int64 label switches the query to the general path (gather + a second
// register sort of the label keys).  Terms are produced position-striped (coalesced log2-table reads, conflict-free
// shared-memory writes); lanes 0 / 1 then run the two strictly sequential fp32 sums cut to cut.
// ---------------------------------------------------------------------------------------------------------------
typedef unsigned long long u64;

__device__ __forceinline__ void cex(u64& a, u64& b) {   // a <- min, b <- max
  const bool sw = b < a;
  const u64 lo = sw ? b : a, hi = sw ? a : b;
  a = lo; b = hi;
}
__device__ __forceinline__ u64 pick(u64 mine, u64 other, bool lower) {   // lower position keeps the smaller key
  const bool other_smaller = other < mine;
  return (other_smaller == lower) ? other : mine;
}
template <int E, int J0>
__device__ __forceinline__ void intra_strides(u64 (&k)[E]) {
#pragma unroll
  for (int j = J0; j > 0; j >>= 1) {
#pragma unroll
    for (int e = 0; e < E; ++e)
      if ((e & j) == 0) cex(k[e], k[e | j]);
  }
}
template <int E>
__device__ __forceinline__ void warp_sort(u64 (&k)[E], int lane) {
  // blocks that fit in one lane
#pragma unroll
  for (int kk = 2; kk <= E; kk <<= 1) {
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int p = e ^ (kk - 1);
      if (p > e) cex(k[e], k[p]);
    }
#pragma unroll
    for (int j = kk >> 2; j > 0; j >>= 1) {
#pragma unroll
      for (int e = 0; e < E; ++e)
        if ((e & j) == 0) cex(k[e], k[e | j]);
    }
  }
  // blocks of m = 2 .. 32 lanes
#pragma unroll 1
  for (int m = 2; m <= 32; m <<= 1) {
    {
      const int mask = m - 1;
      const bool lower = (lane & (m >> 1)) == 0;
      if (E == 1) {
        const u64 o = __shfl_xor_sync(0xffffffffu, k[0], mask);
        k[0] = pick(k[0], o, lower);
      } else {
#pragma unroll
        for (int e = 0; e < E / 2; ++e) {
          const u64 o1 = __shfl_xor_sync(0xffffffffu, k[E - 1 - e], mask);
          const u64 o2 = __shfl_xor_sync(0xffffffffu, k[e], mask);
          k[e] = pick(k[e], o1, lower);
          k[E - 1 - e] = pick(k[E - 1 - e], o2, lower);
        }
      }
    }
#pragma unroll 1
    for (int jl = m >> 2; jl > 0; jl >>= 1) {
      const bool lower = (lane & jl) == 0;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const u64 o = __shfl_xor_sync(0xffffffffu, k[e], jl);
        k[e] = pick(k[e], o, lower);
      }
    }
    intra_strides<E, E / 2>(k);
  }
}

__host__ __device__ constexpr int ndcg_warp_smem(int E) {
  // pay/ti: 33E words (padded), hist8/tp: max(2048, 128E) bytes, hist + hstart + 4 cut arrays: 1024 bytes
  return ((33 * E * 4 + 15) / 16) * 16 + (128 * E > 2048 ? 128 * E : 2048) + 1024;
}

template <int E>
__global__ void __launch_bounds__(128, 4) ndcg_warp_kernel(const float* __restrict__ scores,
                                                        const long long* __restrict__ labels,
                                                        const int* __restrict__ lens, int B, int N, long long ld,
                                                        const long long* __restrict__ ks, int nk,
                                                        const float* __restrict__ log2_table,
                                                        float* __restrict__ ndcg, long long* __restrict__ order) {
  extern __shared__ __align__(16) unsigned char nsm[];
  constexpr int NPAD = 32 * E;
  constexpr int A_BYTES = ((33 * E * 4 + 15) / 16) * 16;
  constexpr int B_BYTES = 128 * E > 2048 ? 128 * E : 2048;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q = blockIdx.x * (blockDim.x >> 5) + warp;
  if (q >= B) return;                                   // whole warps leave; nothing below is block-wide
  unsigned char* base = nsm + (size_t)warp * ndcg_warp_smem(E);
  unsigned int* pay = reinterpret_cast<unsigned int*>(base);            // [33E] sorted (idx << 8 | label), padded
  float* ti = reinterpret_cast<float*>(base);                           // [32E] ideal terms (after pay is consumed)
  unsigned char* hist8 = base + A_BYTES;                                // [64][32] per-lane label counters
  float* tp = reinterpret_cast<float*>(base + A_BYTES);                 // [32E] predicted terms (after hist8)
  int* hist = reinterpret_cast<int*>(base + A_BYTES + B_BYTES);         // [64]
  int* hstart = hist + 64;                                              // [64]
  int* cut_pos = hstart + 64;                                           // [32]
  int* cut_slot = cut_pos + 32;                                         // [32]
  float* cut_p = reinterpret_cast<float*>(cut_slot + 32);               // [32]
  float* cut_i = cut_p + 32;                                            // [32]

  const int n = lens ? min(lens[q], N) : N;
  const float* sq = scores + (long long)q * ld;
  const long long* lq = labels + (long long)q * ld;

  {  // zero the private counters: 2048 bytes = 64 per lane
    uint4* z = reinterpret_cast<uint4*>(hist8) + lane * 4;
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    z[0] = zero; z[1] = zero; z[2] = zero; z[3] = zero;
  }
  __syncwarp();
  u64 k[E];
  bool oor = false;
  {
    // striped (coalesced) loads; the network does not care where an element starts.  All score loads are issued
    // first, then all label loads, and only then is anything consumed: one memory latency covers the whole query
    // (written element by element the label range check serialises the loads: 21 % of the stall samples in
    // profiles/r01_ndcg_full.md)
    float sc[E];
    long long labv[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int i = e * 32 + lane;
      sc[e] = i < n ? sq[i] : 0.f;
    }
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int i = e * 32 + lane;
      labv[e] = i < n ? lq[i] : 0ll;
    }
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int i = e * 32 + lane;
      if (i < n) {
        const long long lab = labv[e];
        const float s = sc[e] + 0.0f;                     // -0.0 -> +0.0: both zeros tie (torch.sort semantics)
        unsigned int u = __float_as_uint(s);
        u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
        u = ~u;                                           // ascending key = descending score
        const bool in_range = lab >= 0 && lab <= 62;
        const unsigned int lb = in_range ? (unsigned int)lab : 0xFFu;
        oor |= !in_range;
        k[e] = ((u64)u << 32) | ((unsigned int)i << 8) | lb;   // ties: lower index first (stable)
        if (in_range) hist8[lb * 32 + lane] += 1;
      } else {
        k[e] = ~0ull;
      }
    }
  }
  const bool fallback = __any_sync(0xffffffffu, oor);
  __syncwarp();
  // counters -> hist[64]: lane handles bins lane and lane + 32 (32 bytes each)
  int cnt0, cnt1;
  {
    const uint4* h = reinterpret_cast<const uint4*>(hist8);
    const uint4 a0 = h[lane * 2], a1 = h[lane * 2 + 1], b0 = h[(lane + 32) * 2], b1 = h[(lane + 32) * 2 + 1];
    cnt0 = __dp4a(a0.x, 0x01010101u, 0u) + __dp4a(a0.y, 0x01010101u, 0u) + __dp4a(a0.z, 0x01010101u, 0u) +
           __dp4a(a0.w, 0x01010101u, 0u) + __dp4a(a1.x, 0x01010101u, 0u) + __dp4a(a1.y, 0x01010101u, 0u) +
           __dp4a(a1.z, 0x01010101u, 0u) + __dp4a(a1.w, 0x01010101u, 0u);
    cnt1 = __dp4a(b0.x, 0x01010101u, 0u) + __dp4a(b0.y, 0x01010101u, 0u) + __dp4a(b0.z, 0x01010101u, 0u) +
           __dp4a(b0.w, 0x01010101u, 0u) + __dp4a(b1.x, 0x01010101u, 0u) + __dp4a(b1.y, 0x01010101u, 0u) +
           __dp4a(b1.z, 0x01010101u, 0u) + __dp4a(b1.w, 0x01010101u, 0u);
    hist[lane] = cnt0; hist[lane + 32] = cnt1;
  }
  __syncwarp();
  {  // first ideal position of each label: descending exclusive scan, bins 62 - 2*lane and 61 - 2*lane per lane
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
  // cuts min(n, k) ranked ascending (ties by slot), one lane per k
  {
    int mine = 0;
    if (lane < nk) {
      const long long kv = ks[lane];
      const long long c = kv < (long long)n ? kv : (long long)n;
      mine = (int)(c < 0 ? 0 : c);
    }
    int rank = 0;
    for (int j = 0; j < nk; ++j) {
      const int other = __shfl_sync(0xffffffffu, mine, j);
      rank += (other < mine || (other == mine && j < lane)) ? 1 : 0;
    }
    if (lane < nk) { cut_pos[rank] = mine; cut_slot[rank] = lane; }
  }

  // One inlined copy of the network: pass 0 sorts the (score, index, label) keys; pass 1 (only when a label is outside
  // [0, 62]) sorts the raw int64 label keys for the ideal ordering.
  const int passes = fallback ? 2 : 1;
#pragma unroll 1
  for (int pass = 0; pass < passes; ++pass) {
    if (pass == 1) {
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const int i = e * 32 + lane;
        k[e] = i < n ? label_key(lq[i]) : ~0ull;
      }
    }
    warp_sort<E>(k, lane);
    if (pass == 1) {
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const int p = lane * E + e;
        if (p < n) ti[p] = gain_of(label_from_key(k[e])) / log2_table[p];
      }
      break;
    }
    // sorted payloads -> shared memory (blocked positions, one pad word per 32 keeps the banks distinct)
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int p = lane * E + e;
      pay[p + (p >> 5)] = (unsigned int)(k[e] & 0xFFFFFFFFull);
    }
    __syncwarp();
    // predicted terms, position-striped.  tp overwrites the byte counters (already reduced).
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int i = e * 32 + lane;
      if (i < n) {
        const unsigned int w = pay[i + (i >> 5)];
        const unsigned int idx = w >> 8;
        if (order != nullptr) order[(long long)q * ld + i] = (long long)idx;
        const long long lab = fallback ? lq[idx] : (long long)(w & 0xFFu);
        // a zero numerator sends the IEEE division to its slow path (FCHK); 0 / x is +0 for every x in the table
        const float g = gain_of(lab);
        const float t = (g == 0.f ? 1.0f : g) / log2_table[i];
        tp[i] = g == 0.f ? 0.f : t;
      }
    }
    __syncwarp();                                          // pay fully consumed: ti may overwrite it
    if (!fallback) {
      // ideal order = labels descending: label L fills positions [hstart[L], hstart[L] + hist[L])
      const unsigned int ne_lo = __ballot_sync(0xffffffffu, cnt0 > 0), ne_hi = __ballot_sync(0xffffffffu, cnt1 > 0);
      u64 present = ((u64)ne_hi << 32) | ne_lo;
      while (present) {
        const int L = 63 - __clzll((long long)present);
        present &= ~(1ull << L);
        const int s0 = hstart[L], c = hist[L];
        const float g = gain_of((long long)L);
        if (g == 0.f) {
          for (int i = s0 + lane; i < s0 + c; i += 32) ti[i] = 0.f;
        } else {
#pragma unroll 4
          for (int i = s0 + lane; i < s0 + c; i += 32) ti[i] = g / log2_table[i];
        }
      }
    }
  }
  __syncwarp();
  // strictly sequential fp32 sums, walked cut to cut: lane 0 predicted, lane 1 ideal
  if (lane < 2) {
    const float* t = lane == 0 ? tp : ti;
    float* outc = lane == 0 ? cut_p : cut_i;
    float acc = 0.f;
    int i = 0;
    for (int j = 0; j < nk; ++j) {
      const int c = cut_pos[j];
      for (; i < c && (i & 3) != 0; ++i) acc = acc + t[i];
      for (; i + 8 <= c; i += 8) {
        const float4 v0 = *reinterpret_cast<const float4*>(t + i), v1 = *reinterpret_cast<const float4*>(t + i + 4);
        acc = acc + v0.x; acc = acc + v0.y; acc = acc + v0.z; acc = acc + v0.w;
        acc = acc + v1.x; acc = acc + v1.y; acc = acc + v1.z; acc = acc + v1.w;
      }
      for (; i < c; ++i) acc = acc + t[i];
      outc[cut_slot[j]] = acc;
    }
  }
  __syncwarp();
  if (lane < nk) {
    const float p = cut_p[lane], t = cut_i[lane];
    ndcg[(long long)q * nk + lane] = (t <= 1e-6f) ? 1.0f : p / t;
  }
  (void)NPAD;
}


// ---------------------------------------------------------------------------------------------------------------
// Warp-per-query path, 32-bit keys (round 2; the default for N <= 1024).
// The 64-bit kernel above spends its time in compare-exchanges of u64 keys (2 ISETP + 4 SEL each on the half-rate
// integer pipe: ALU pipe 50 %, 14.8 k warp instructions per 1024-key query, 128 registers; profiles/r01_ndcg_full.md).
// Here the network sorts ONE 32-bit word per element, key32 = [top 22 bits of the descending score key | index(10)],
// so an in-register compare-exchange is 2 IMNMX and a shuffle stage is SHFL + 1 predicated IMNMX per element, with
// half the registers.  Dropping the low 10 score bits only mis-orders elements whose keys agree in the top 22 bits
// (they come out in index order instead of low-bit order); the exact order is restored by
//   1. gathering  sub = [low 10 score bits | index | label]  for every sorted element from a shared-memory side table,
//   2. odd-even transposition rounds that swap ADJACENT elements iff their top-22 bits agree and sub is out of order
//      (runs of colliding keys are 2-3 long for real score distributions: one swapping round + one confirming round;
//      skipped entirely when no adjacent top-22 collision exists),
//   3. a bail-out after 3 rounds (pathological inputs: hundreds of scores inside a 2^-13 relative interval) to a
//      shared-memory bitonic sort of the full 64-bit keys -- also the sort the general-label path uses.
// The result is the same total order as the 64-bit kernel: (score descending, index ascending), bit for bit.
// Labels ride in `sub`, the ideal ordering comes from the byte-counter histogram exactly as above.
// ---------------------------------------------------------------------------------------------------------------
typedef unsigned int u32;

__device__ __forceinline__ void cex32(u32& a, u32& b) {   // a <- min, b <- max
  const u32 lo = min(a, b), hi = max(a, b);
  a = lo; b = hi;
}
template <int E>
__device__ __forceinline__ void warp_sort32(u32 (&k)[E], int lane) {
#pragma unroll
  for (int kk = 2; kk <= E; kk <<= 1) {
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int p = e ^ (kk - 1);
      if (p > e) cex32(k[e], k[p]);
    }
#pragma unroll
    for (int j = kk >> 2; j > 0; j >>= 1) {
#pragma unroll
      for (int e = 0; e < E; ++e)
        if ((e & j) == 0) cex32(k[e], k[e | j]);
    }
  }
#pragma unroll 1
  for (int m = 2; m <= 32; m <<= 1) {
    {
      const int mask = m - 1;
      const bool lower = (lane & (m >> 1)) == 0;
      if (E == 1) {
        const u32 o = __shfl_xor_sync(0xffffffffu, k[0], mask);
        k[0] = lower ? min(k[0], o) : max(k[0], o);
      } else {
#pragma unroll
        for (int e = 0; e < E / 2; ++e) {
          const u32 o1 = __shfl_xor_sync(0xffffffffu, k[E - 1 - e], mask);
          const u32 o2 = __shfl_xor_sync(0xffffffffu, k[e], mask);
          k[e] = lower ? min(k[e], o1) : max(k[e], o1);
          k[E - 1 - e] = lower ? min(k[E - 1 - e], o2) : max(k[E - 1 - e], o2);
        }
      }
    }
#pragma unroll 1
    for (int jl = m >> 2; jl > 0; jl >>= 1) {
      const bool lower = (lane & jl) == 0;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const u32 o = __shfl_xor_sync(0xffffffffu, k[e], jl);
        k[e] = lower ? min(k[e], o) : max(k[e], o);
      }
    }
#pragma unroll
    for (int j = E / 2; j > 0; j >>= 1) {
#pragma unroll
      for (int e = 0; e < E; ++e)
        if ((e & j) == 0) cex32(k[e], k[e | j]);
    }
  }
}

// Warp-level bitonic sort of npad u64 keys in shared memory (ascending).  Slow path only: general int64 labels and
// the repair bail-out.
__device__ __noinline__ void warp_smem_sort_u64(u64* keys, int npad, int lane) {
  for (int kk = 2; kk <= npad; kk <<= 1) {
    for (int j = kk >> 1; j > 0; j >>= 1) {
      for (int t = lane; t < (npad >> 1); t += 32) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int ixj = i | j;
        const u64 a = keys[i], b = keys[ixj];
        const bool up = (i & kk) == 0;
        if ((a > b) == up) { keys[i] = b; keys[ixj] = a; }
      }
      __syncwarp();
    }
  }
}

// swap (ka, sa) <-> (kb, sb) iff the top-22 bits agree and sub is out of order; returns whether it swapped
__device__ __forceinline__ bool fix_pair(u32& ka, u32& sa, u32& kb, u32& sb) {
  const bool sw = ((ka ^ kb) < 1024u) && (sa > sb);
  const u32 k0 = sw ? kb : ka, k1 = sw ? ka : kb, s0 = sw ? sb : sa, s1 = sw ? sa : sb;
  ka = k0; kb = k1; sa = s0; sb = s1;
  return sw;
}

__host__ __device__ constexpr int ndcg_warp32_a_bytes(int E) { return ((33 * E * 4 + 15) / 16) * 16; }
__host__ __device__ constexpr int ndcg_warp32_b_bytes(int E) { return 128 * E > 2048 ? 128 * E : 2048; }
__host__ __device__ constexpr int ndcg_warp32_smem(int E) {
  // A: sorted payloads [33E] words | B: byte counters -> predicted terms | C: side table -> ideal terms | 1 KB of small
  // arrays.  A and B are contiguous: together they hold the 32E u64 keys of the slow-path sort.
  return ndcg_warp32_a_bytes(E) + ndcg_warp32_b_bytes(E) + 128 * E + 1024;
}

template <int E>
__global__ void __launch_bounds__(128, 4) ndcg_warp32_kernel(const float* __restrict__ scores,
                                                          const long long* __restrict__ labels,
                                                          const int* __restrict__ lens, int B, int N, long long ld,
                                                          const long long* __restrict__ ks, int nk,
                                                          const float* __restrict__ log2_table,
                                                          float* __restrict__ ndcg, long long* __restrict__ order) {
  extern __shared__ __align__(16) unsigned char nsm[];
  constexpr int NPAD = 32 * E;
  constexpr int A_BYTES = ndcg_warp32_a_bytes(E);
  constexpr int B_BYTES = ndcg_warp32_b_bytes(E);
  static_assert(A_BYTES + B_BYTES >= 8 * NPAD, "slow-path sort needs 8 bytes per key in A|B");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q = blockIdx.x * (blockDim.x >> 5) + warp;
  if (q >= B) return;                                   // whole warps leave; nothing below is block-wide
  unsigned char* base = nsm + (size_t)warp * ndcg_warp32_smem(E);
  u32* pay = reinterpret_cast<u32*>(base);                              // A: [33E] sorted (idx << 8 | label), padded
  u64* slow = reinterpret_cast<u64*>(base);                             // A|B: [32E] u64 keys (slow paths)
  unsigned char* hist8 = base + A_BYTES;                                // B: [64][32] per-lane label counters
  float* tp = reinterpret_cast<float*>(base + A_BYTES);                 // B: [32E] predicted terms (after hist8)
  u32* side = reinterpret_cast<u32*>(base + A_BYTES + B_BYTES);         // C: [32E] (low10 << 18 | idx << 8 | label)
  float* ti = reinterpret_cast<float*>(base + A_BYTES + B_BYTES);       // C: [32E] ideal terms (after side)
  int* hist = reinterpret_cast<int*>(base + A_BYTES + B_BYTES + 128 * E);   // [64]
  int* hstart = hist + 64;                                              // [64]
  int* cut_pos = hstart + 64;                                           // [32]
  int* cut_slot = cut_pos + 32;                                         // [32]
  float* cut_p = reinterpret_cast<float*>(cut_slot + 32);               // [32]
  float* cut_i = cut_p + 32;                                            // [32]

  const int n = lens ? min(lens[q], N) : N;
  const float* sq = scores + (long long)q * ld;
  const long long* lq = labels + (long long)q * ld;

  {
    uint4* z = reinterpret_cast<uint4*>(hist8) + lane * 4;
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    z[0] = zero; z[1] = zero; z[2] = zero; z[3] = zero;
  }
  __syncwarp();
  u32 k[E];
  bool oor = false;
  {
    float sc[E];
    long long labv[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int i = e * 32 + lane;
      sc[e] = i < n ? sq[i] : 0.f;
    }
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int i = e * 32 + lane;
      labv[e] = i < n ? lq[i] : 0ll;
    }
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int i = e * 32 + lane;
      if (i < n) {
        const long long lab = labv[e];
        const float s = sc[e] + 0.0f;                     // -0.0 -> +0.0: both zeros tie (torch.sort semantics)
        u32 u = __float_as_uint(s);
        u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
        u = ~u;                                           // ascending key = descending score
        const bool in_range = lab >= 0 && lab <= 62;
        const u32 lb = in_range ? (u32)lab : 0xFFu;
        oor |= !in_range;
        k[e] = (u & 0xFFFFFC00u) | (u32)i;                // ties in the top 22 bits: lower index first
        side[i] = ((u & 0x3FFu) << 18) | ((u32)i << 8) | lb;
        if (in_range) hist8[lb * 32 + lane] += 1;
      } else {
        k[e] = 0xFFFFFFFFu;
        side[i] = 0xFFFFFFFFu;                            // pads compare as "in order" with anything real
      }
    }
  }
  const bool fallback = __any_sync(0xffffffffu, oor);
  __syncwarp();
  int cnt0, cnt1;
  {
    const uint4* h = reinterpret_cast<const uint4*>(hist8);
    const uint4 a0 = h[lane * 2], a1 = h[lane * 2 + 1], b0 = h[(lane + 32) * 2], b1 = h[(lane + 32) * 2 + 1];
    cnt0 = __dp4a(a0.x, 0x01010101u, 0u) + __dp4a(a0.y, 0x01010101u, 0u) + __dp4a(a0.z, 0x01010101u, 0u) +
           __dp4a(a0.w, 0x01010101u, 0u) + __dp4a(a1.x, 0x01010101u, 0u) + __dp4a(a1.y, 0x01010101u, 0u) +
           __dp4a(a1.z, 0x01010101u, 0u) + __dp4a(a1.w, 0x01010101u, 0u);
    cnt1 = __dp4a(b0.x, 0x01010101u, 0u) + __dp4a(b0.y, 0x01010101u, 0u) + __dp4a(b0.z, 0x01010101u, 0u) +
           __dp4a(b0.w, 0x01010101u, 0u) + __dp4a(b1.x, 0x01010101u, 0u) + __dp4a(b1.y, 0x01010101u, 0u) +
           __dp4a(b1.z, 0x01010101u, 0u) + __dp4a(b1.w, 0x01010101u, 0u);
    hist[lane] = cnt0; hist[lane + 32] = cnt1;
  }
  __syncwarp();
  {
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
  {
    int mine = 0;
    if (lane < nk) {
      const long long kv = ks[lane];
      const long long c = kv < (long long)n ? kv : (long long)n;
      mine = (int)(c < 0 ? 0 : c);
    }
    int rank = 0;
    for (int j = 0; j < nk; ++j) {
      const int other = __shfl_sync(0xffffffffu, mine, j);
      rank += (other < mine || (other == mine && j < lane)) ? 1 : 0;
    }
    if (lane < nk) { cut_pos[rank] = mine; cut_slot[rank] = lane; }
  }

  // ---- sort by the truncated keys, then restore the exact order -------------------------------------------
  warp_sort32<E>(k, lane);
  u32 sub[E];
#pragma unroll
  for (int e = 0; e < E; ++e) sub[e] = k[e] == 0xFFFFFFFFu ? 0xFFFFFFFFu : side[k[e] & 0x3FFu];
  // any adjacent pair with equal top-22 bits?  (blocked positions p = lane * E + e; the pair across a lane boundary
  // is (E-1 of lane, 0 of lane + 1))
  const u32 k_next0 = __shfl_down_sync(0xffffffffu, k[0], 1);
  bool coll = lane < 31 && ((k[E - 1] ^ k_next0) < 1024u) && k_next0 != 0xFFFFFFFFu;
#pragma unroll
  for (int e = 0; e + 1 < E; ++e) coll |= ((k[e] ^ k[e + 1]) < 1024u) && k[e + 1] != 0xFFFFFFFFu;
  if (__any_sync(0xffffffffu, coll)) {
    int rounds = 0;
    bool again = true;
    while (again && rounds < 3) {
      bool sw = false;
      if (E > 1) {
        // even phase: pairs (e, e + 1), e even -- all inside a lane
#pragma unroll
        for (int e = 0; e + 1 < E; e += 2) sw |= fix_pair(k[e], sub[e], k[e + 1], sub[e + 1]);
        // odd phase: pairs (e, e + 1), e odd, and the lane-boundary pair
#pragma unroll
        for (int e = 1; e + 1 < E; e += 2) sw |= fix_pair(k[e], sub[e], k[e + 1], sub[e + 1]);
        {
          u32 nk0 = __shfl_down_sync(0xffffffffu, k[0], 1), ns0 = __shfl_down_sync(0xffffffffu, sub[0], 1);
          u32 pk = __shfl_up_sync(0xffffffffu, k[E - 1], 1), ps = __shfl_up_sync(0xffffffffu, sub[E - 1], 1);
          if (lane < 31) {      // I hold the lower element of (my E-1, next lane's 0)
            u32 a = k[E - 1], b = sub[E - 1];
            sw |= fix_pair(a, b, nk0, ns0);
            k[E - 1] = a; sub[E - 1] = b;
          }
          if (lane > 0) {       // I hold the upper element of (previous lane's E-1, my 0): same decision, other half
            u32 a = k[0], b = sub[0];
            fix_pair(pk, ps, a, b);
            k[0] = a; sub[0] = b;
          }
        }
      } else {
        // one element per lane: even phase pairs lanes (2j, 2j+1), odd phase (2j+1, 2j+2)
#pragma unroll
        for (int phase = 0; phase < 2; ++phase) {
          const bool lower = ((lane ^ phase) & 1) == 0;
          const int partner = lower ? lane + 1 : lane - 1;
          const bool valid = partner >= 0 && partner < 32;
          const u32 ok = __shfl_sync(0xffffffffu, k[0], valid ? partner : lane);
          const u32 os = __shfl_sync(0xffffffffu, sub[0], valid ? partner : lane);
          if (valid) {
            u32 a = k[0], b = sub[0], c = ok, d = os;
            if (lower) { sw |= fix_pair(a, b, c, d); k[0] = a; sub[0] = b; }
            else { fix_pair(c, d, a, b); k[0] = a; sub[0] = b; }
          }
        }
      }
      again = __any_sync(0xffffffffu, sw);
      ++rounds;
    }
    if (again) {
      // still swapping after 3 rounds: long runs of near-equal scores.  Exact fallback: sort the FULL keys.
      __syncwarp();
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const int p = lane * E + e;
        const u32 u = (k[e] & 0xFFFFFC00u) | (sub[e] >> 18);
        slow[p] = k[e] == 0xFFFFFFFFu ? ~0ull : (((u64)u << 32) | (u64)(sub[e] & 0x3FFFFu));
      }
      __syncwarp();
      warp_smem_sort_u64(slow, NPAD, lane);
#pragma unroll
      for (int e = 0; e < E; ++e) sub[e] = (u32)(slow[lane * E + e] & 0xFFFFFFFFull);   // low 18 bits: idx | label
      __syncwarp();
    }
  }
  // side table consumed (sub holds idx | label per sorted position): region C may now hold the ideal terms
  __syncwarp();
  if (fallback) {
    // general int64 labels: ideal ordering = labels sorted descending (full 64-bit keys, shared-memory sort in A|B)
    for (int i = lane; i < NPAD; i += 32) slow[i] = i < n ? label_key(lq[i]) : ~0ull;
    __syncwarp();
    warp_smem_sort_u64(slow, NPAD, lane);
    for (int p = lane; p < n; p += 32) ti[p] = gain_of(label_from_key(slow[p])) / log2_table[p];
    __syncwarp();
  }
  // sorted payloads -> shared memory (blocked positions, one pad word per 32 keeps the banks distinct)
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int p = lane * E + e;
    pay[p + (p >> 5)] = sub[e] & 0x3FFFFu;
  }
  __syncwarp();
  // predicted terms, position-striped.  tp overwrites the byte counters (already reduced).
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int i = e * 32 + lane;
    if (i < n) {
      const u32 w = pay[i + (i >> 5)];
      const u32 idx = w >> 8;
      if (order != nullptr) order[(long long)q * ld + i] = (long long)idx;
      const long long lab = fallback ? lq[idx] : (long long)(w & 0xFFu);
      const float g = gain_of(lab);
      const float t = (g == 0.f ? 1.0f : g) / log2_table[i];   // 0 / x takes the slow IEEE path; it is +0 anyway
      tp[i] = g == 0.f ? 0.f : t;
    }
  }
  __syncwarp();
  if (!fallback) {
    const unsigned int ne_lo = __ballot_sync(0xffffffffu, cnt0 > 0), ne_hi = __ballot_sync(0xffffffffu, cnt1 > 0);
    u64 present = ((u64)ne_hi << 32) | ne_lo;
    while (present) {
      const int L = 63 - __clzll((long long)present);
      present &= ~(1ull << L);
      const int s0 = hstart[L], c = hist[L];
      const float g = gain_of((long long)L);
      if (g == 0.f) {
        for (int i = s0 + lane; i < s0 + c; i += 32) ti[i] = 0.f;
      } else {
#pragma unroll 4
        for (int i = s0 + lane; i < s0 + c; i += 32) ti[i] = g / log2_table[i];
      }
    }
  }
  __syncwarp();
  if (lane < 2) {
    const float* t = lane == 0 ? tp : ti;
    float* outc = lane == 0 ? cut_p : cut_i;
    float acc = 0.f;
    int i = 0;
    for (int j = 0; j < nk; ++j) {
      const int c = cut_pos[j];
      for (; i < c && (i & 3) != 0; ++i) acc = acc + t[i];
      for (; i + 8 <= c; i += 8) {
        const float4 v0 = *reinterpret_cast<const float4*>(t + i), v1 = *reinterpret_cast<const float4*>(t + i + 4);
        acc = acc + v0.x; acc = acc + v0.y; acc = acc + v0.z; acc = acc + v0.w;
        acc = acc + v1.x; acc = acc + v1.y; acc = acc + v1.z; acc = acc + v1.w;
      }
      for (; i < c; ++i) acc = acc + t[i];
      outc[cut_slot[j]] = acc;
    }
  }
  __syncwarp();
  if (lane < nk) {
    const float p = cut_p[lane], t = cut_i[lane];
    ndcg[(long long)q * nk + lane] = (t <= 1e-6f) ? 1.0f : p / t;
  }
}

template <int E>
static int launch_ndcg_warp32(const float* scores, const long long* labels, const int* lens, int B, int N, long long ld,
                              const long long* ks, int nk, const float* log2_table, float* ndcg, long long* order,
                              cudaStream_t s) {
  int wpb = B / 148;                      // fill the 148 SMs first, then stack warps per block
  wpb = wpb < 1 ? 1 : (wpb > 4 ? 4 : wpb);
  if (wpb == 3) wpb = 2;
  const size_t smem = (size_t)wpb * ndcg_warp32_smem(E);
  static size_t configured = 48 * 1024;
  if (smem > configured) {
    if (cudaFuncSetAttribute(ndcg_warp32_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
        cudaSuccess)
      return LR2_ERR_CUDA;
    configured = smem;
  }
  ndcg_warp32_kernel<E><<<(B + wpb - 1) / wpb, wpb * 32, smem, s>>>(scores, labels, lens, B, N, ld, ks, nk,
                                                                   log2_table, ndcg, order); LR2_LAUNCHED(1);
  LR2_RETURN_LAUNCH();
}

template <int E>
static int launch_ndcg_warp(const float* scores, const long long* labels, const int* lens, int B, int N, long long ld,
                            const long long* ks, int nk, const float* log2_table, float* ndcg, long long* order,
                            cudaStream_t s) {
  int wpb = B / 148;                      // fill the 148 SMs first, then stack warps per block
  wpb = wpb < 1 ? 1 : (wpb > 4 ? 4 : wpb);
  if (wpb == 3) wpb = 2;
  const size_t smem = (size_t)wpb * ndcg_warp_smem(E);
  static size_t configured = 48 * 1024;
  if (smem > configured) {
    if (cudaFuncSetAttribute(ndcg_warp_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
        cudaSuccess)
      return LR2_ERR_CUDA;
    configured = smem;
  }
  ndcg_warp_kernel<E><<<(B + wpb - 1) / wpb, wpb * 32, smem, s>>>(scores, labels, lens, B, N, ld, ks, nk, log2_table,
                                                                 ndcg, order); LR2_LAUNCHED(1);
  LR2_RETURN_LAUNCH();
}

}  // namespace lr2

using namespace lr2;

extern "C" int lr2_ndcg_at_k(const float* scores, const long long* labels, const int* lens, int B, int N, long long ld,
                             const long long* ks, int nk, const float* log2_table, float* ndcg, long long* order,
                             void* stream) {
  if (B <= 0 || N <= 0 || nk <= 0 || ld < N) return LR2_ERR_BAD_SHAPE;
  if (N > 4096 || nk > lr2::NDCG_MAX_K) return LR2_ERR_UNSUPPORTED;
  {
    static const int legacy = [] { const char* e = getenv("LR2_NDCG_LEGACY"); return e && e[0] == '1' ? 1 : 0; }();
    // one warp per query wins whenever there are enough queries to fill the SMs (or the query is short); a few long
    // queries are latency-bound in a single warp and keep the block-per-query kernel (profiles/r01_ndcg_sweep.md)
    static const int force_warp = [] { const char* e = getenv("LR2_NDCG_WARP"); return e && e[0] == '1' ? 1 : 0; }();
    static const int warp64 = [] { const char* e = getenv("LR2_NDCG_WARP64"); return e && e[0] == '1' ? 1 : 0; }();
    if (!legacy && !warp64 && N <= 1024 && (force_warp || N <= 256 || B >= 1024)) {
      cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
#define LR2_NDCG_WARP32(E_) return launch_ndcg_warp32<E_>(scores, labels, lens, B, N, ld, ks, nk, log2_table, ndcg, order, s)
      if (N <= 32) LR2_NDCG_WARP32(1);
      if (N <= 64) LR2_NDCG_WARP32(2);
      if (N <= 128) LR2_NDCG_WARP32(4);
      if (N <= 256) LR2_NDCG_WARP32(8);
      if (N <= 512) LR2_NDCG_WARP32(16);
      LR2_NDCG_WARP32(32);
#undef LR2_NDCG_WARP32
    }
    if (!legacy && N <= 1024 && (force_warp || N <= 256 || B >= 1024)) {
      cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
#define LR2_NDCG_WARP(E_) return launch_ndcg_warp<E_>(scores, labels, lens, B, N, ld, ks, nk, log2_table, ndcg, order, s)
      if (N <= 32) LR2_NDCG_WARP(1);
      if (N <= 64) LR2_NDCG_WARP(2);
      if (N <= 128) LR2_NDCG_WARP(4);
      if (N <= 256) LR2_NDCG_WARP(8);
      if (N <= 512) LR2_NDCG_WARP(16);
      LR2_NDCG_WARP(32);
#undef LR2_NDCG_WARP
    }
  }
  int npad = 2;
  while (npad < N) npad <<= 1;
  const size_t smem = (size_t)npad * (8 + 8 + 4 + 4);
  static size_t configured = 40 * 1024;   // static shared memory of the kernel counts against the 48 KB default
  if (smem > configured) {
    if (cudaFuncSetAttribute(ndcg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return LR2_ERR_CUDA;
    configured = smem;
  }
  static const int pair_index = [] { const char* e = getenv("LR2_NDCG_BLOCK_PAIRS"); return e && e[0] == '1' ? 1 : 0; }();
  int threads = npad / 2;
  if (threads < 32) threads = 64;
  if (threads > 512) threads = 512;
  if (threads < 64) threads = 64;
  ndcg_kernel<<<B, threads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(scores, labels, lens, N, ld, ks, nk,
                                                                            log2_table, ndcg, order, npad, pair_index); LR2_LAUNCHED(1);
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
