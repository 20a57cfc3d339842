// PPO / ranking row kernels (memory-bound, tiny): policy loss fwd+bwd, clipped value loss,
// pairwise hinge, SmoothL1, rollout argsort + permutation compose, Plackett-Luce sampler,
// GAE reverse scan.  Arithmetic follows SURVEY.md §8(a') line by line; the CPU restatement
// that checks these kernels is oracle/ppo_rows.c.
#include "common.cuh"
#include "det_math.h"

namespace lr2 {

constexpr int PL_THREADS = 256;

__device__ __forceinline__ float clamp_log(float t) { return logf(fmaxf(t, 1e-20f)); }  // ref: finetune/ppo.py:431-432

// deterministic block sum (fixed tree) of up to 4 values per thread
template <int NV>
__device__ __forceinline__ void block_sum(float (&v)[NV], float* sm /* [NV][PL_THREADS/32] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i] = warp_sum(v[i]);
    if (lane == 0) sm[i * (PL_THREADS / 32) + warp] = v[i];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < PL_THREADS / 32; ++w) s += sm[i * (PL_THREADS / 32) + w];
    v[i] = s;
  }
  __syncthreads();
}

// ref: finetune/ppo.py:38-55, 544-575
__global__ void __launch_bounds__(PL_THREADS)
ppo_policy_loss_kernel(const float* __restrict__ s, const float* __restrict__ s_old,
                       const float* __restrict__ reward, const float* __restrict__ v_old,
                       const long long* __restrict__ pi, int B, int n, int kp, float w_kl, float w_ent, float margin,
                       float adv_eps, float* __restrict__ out_scalars, float* __restrict__ kl_out,
                       float* __restrict__ ent_out, float* __restrict__ radj_out, float* __restrict__ adv_out,
                       float* __restrict__ ds) {
  __shared__ float sm[4 * (PL_THREADS / 32)];
  __shared__ int bad_index;
  if (threadIdx.x == 0) bad_index = 0;
  __syncthreads();
  float acc[4] = {0.f, 0.f, 0.f, 0.f};  // sum h, cnt, sum|A|, sum H
  // pi is [B, kp]: the reference's RankLoss takes any [B, k] index list (finetune/ppo.py:43-46); the update passes
  // next_state[:, -2:] (k = 2).  An index outside [0, n) never dereferences s: the row is skipped and the loss is NaN.
  for (int b = threadIdx.x; b < B; b += PL_THREADS) {
    const long long* pchk = pi + (long long)b * kp;
    for (int i = 0; i < kp; ++i)
      if (pchk[i] < 0 || pchk[i] >= n) bad_index = 1;
  }
  __syncthreads();
  const bool bad = bad_index != 0;
  for (int b = threadIdx.x; b < B; b += PL_THREADS) {
    const float* sb = s + (long long)b * n;
    const float* so = s_old + (long long)b * n;
    float mx = -INFINITY, mo = -INFINITY;
    for (int j = 0; j < n; ++j) { mx = fmaxf(mx, sb[j]); mo = fmaxf(mo, so[j]); }
    float den = 0.f, deo = 0.f;
    for (int j = 0; j < n; ++j) { den += expf(sb[j] - mx); deo += expf(so[j] - mo); }
    float kl = 0.f, H = 0.f;
    for (int j = 0; j < n; ++j) {
      const float p = expf(sb[j] - mx) / den, po = expf(so[j] - mo) / deo;
      const float lp = clamp_log(p);
      kl += po * (clamp_log(po) - lp);
      H -= p * lp;
    }
    const float radj = reward[b] - kl * w_kl;
    const float A = radj - v_old[b];
    const bool flip = !(A >= adv_eps);
    const long long* pb = pi + (long long)b * kp;
    float hs = 0.f, hc = 0.f;
    for (int i = 0; i < kp && !bad; ++i) {
      const float si = sb[flip ? pb[kp - 1 - i] : pb[i]];
      for (int j = i + 1; j < kp; ++j) {
        const float sj = sb[flip ? pb[kp - 1 - j] : pb[j]];
        const float h = fmaxf(margin - (si - sj), 0.f);
        hs += h;
        hc += (h > 0.f) ? 1.f : 0.f;
      }
    }
    kl_out[b] = kl; ent_out[b] = H; radj_out[b] = radj; adv_out[b] = A;
    acc[0] += hs; acc[1] += hc; acc[2] += fabsf(A); acc[3] += H;
  }
  block_sum<4>(acc, sm);
  const float sum_h = acc[0], cnt = acc[1], sum_abs = acc[2], sum_H = acc[3];
  const float L_rank = (cnt == 0.f) ? sum_h : sum_h / cnt;
  const float invB = 1.f / (float)B;
  if (threadIdx.x == 0) {
    // mean_b(L_rank*|A_b| - w_e*H_b)
    out_scalars[0] = bad ? __int_as_float(0x7fc00000) : (L_rank * sum_abs - w_ent * sum_H) * invB;
    out_scalars[1] = bad ? __int_as_float(0x7fc00000) : L_rank;
    out_scalars[2] = cnt;
    out_scalars[3] = sum_abs;
  }
  if (ds == nullptr) return;
  const float hinge_coef = (cnt == 0.f) ? 0.f : sum_abs * invB / cnt;
  for (int b = threadIdx.x; b < B; b += PL_THREADS) {
    const float* sb = s + (long long)b * n;
    const float* so = s_old + (long long)b * n;
    float* db = ds + (long long)b * n;
    float mx = -INFINITY, mo = -INFINITY;
    for (int j = 0; j < n; ++j) { mx = fmaxf(mx, sb[j]); mo = fmaxf(mo, so[j]); }
    float den = 0.f, deo = 0.f;
    for (int j = 0; j < n; ++j) { den += expf(sb[j] - mx); deo += expf(so[j] - mo); }
    const float A = adv_out[b];
    const float sgn = (A > 0.f) ? 1.f : ((A < 0.f) ? -1.f : 0.f);
    // dKL/ds_k = -sum_j po_j m_j (delta_jk - p_k);  dH/ds_k = p_k (g_k - sum_j g_j p_j), g_j = -(lg p_j + m_j)
    float sum_pom = 0.f, sum_gp = 0.f;
    for (int j = 0; j < n; ++j) {
      const float p = expf(sb[j] - mx) / den, po = expf(so[j] - mo) / deo;
      const float mj = (p >= 1e-20f) ? 1.f : 0.f;
      sum_pom += po * mj;
      sum_gp += -(clamp_log(p) + mj) * p;
    }
    for (int k = 0; k < n; ++k) {
      const float p = expf(sb[k] - mx) / den, po = expf(so[k] - mo) / deo;
      const float mk = (p >= 1e-20f) ? 1.f : 0.f;
      const float dkl = -(po * mk - p * sum_pom);
      const float dH = p * (-(clamp_log(p) + mk) - sum_gp);
      db[k] = invB * L_rank * sgn * (-w_kl) * dkl - w_ent * invB * dH;
    }
    const bool flip = !(A >= adv_eps);
    const long long* pb = pi + (long long)b * kp;
    for (int i = 0; i < kp && !bad; ++i) {
      const long long oi = flip ? pb[kp - 1 - i] : pb[i];
      for (int j = i + 1; j < kp; ++j) {
        const long long oj = flip ? pb[kp - 1 - j] : pb[j];
        if (margin - (sb[oi] - sb[oj]) > 0.f) { db[oi] -= hinge_coef; db[oj] += hinge_coef; }
      }
    }
  }
}

// ref: finetune/ppo.py:494-498
__global__ void __launch_bounds__(PL_THREADS)
clipped_value_loss_kernel(const float* __restrict__ v, const float* __restrict__ ret, const float* __restrict__ v_old,
                          int B, float clip, float* __restrict__ out_loss, float* __restrict__ dv) {
  __shared__ float sm[PL_THREADS / 32];
  float acc[1] = {0.f};
  const float invB = 1.f / (float)B;
  for (int b = threadIdx.x; b < B; b += PL_THREADS) {
    const float d = v[b] - v_old[b];
    const float dc = fminf(fmaxf(d, -clip), clip);
    const float vc = v_old[b] + dc;
    const float e1 = vc - ret[b], e2 = v[b] - ret[b];
    const float l1 = e1 * e1, l2 = e2 * e2;
    acc[0] += fmaxf(l1, l2);
    if (dv != nullptr) {
      const float inside = (d >= -clip && d <= clip) ? 1.f : 0.f;
      const float g1 = 2.f * e1 * inside, g2 = 2.f * e2;
      float g;
      if (l1 > l2) g = g1;
      else if (l1 < l2) g = g2;
      else g = 0.5f * g1 + 0.5f * g2;  // torch.maximum splits ties evenly
      dv[b] = g * invB;
    }
  }
  block_sum<1>(acc, sm);
  if (threadIdx.x == 0) out_loss[0] = acc[0] * invB;
}

// ref: finetune/reward_pair_dataloader.py:355-358
__global__ void __launch_bounds__(PL_THREADS)
pair_hinge_kernel(const float* __restrict__ c, const float* __restrict__ r, int B, float margin,
                  float* __restrict__ out, float* __restrict__ dc, float* __restrict__ dr) {
  __shared__ float sm[2 * (PL_THREADS / 32)];
  float acc[2] = {0.f, 0.f};
  const float invB = 1.f / (float)B;
  for (int b = threadIdx.x; b < B; b += PL_THREADS) {
    const float h = margin - (c[b] - r[b]);
    acc[0] += fmaxf(h, 0.f);
    acc[1] += (c[b] > r[b]) ? 1.f : 0.f;
    const float g = (h > 0.f) ? invB : 0.f;
    if (dc != nullptr) dc[b] = -g;
    if (dr != nullptr) dr[b] = g;
  }
  block_sum<2>(acc, sm);
  if (threadIdx.x == 0) { out[0] = acc[0] * invB; out[1] = acc[1] * invB; }
}

// ref: finetune/pointwise.py:229
__global__ void __launch_bounds__(PL_THREADS)
smooth_l1_kernel(const float* __restrict__ x, const long long* __restrict__ tgt, long long n, float beta,
                 float* __restrict__ out_loss, float* __restrict__ dx) {
  __shared__ float sm[PL_THREADS / 32];
  float acc[1] = {0.f};
  const float invn = 1.f / (float)n;
  for (long long i = threadIdx.x; i < n; i += PL_THREADS) {
    const float d = x[i] - (float)tgt[i];
    const float ad = fabsf(d);
    acc[0] += (ad < beta) ? 0.5f * d * d / beta : ad - 0.5f * beta;
    if (dx != nullptr) dx[i] = ((ad < beta) ? d / beta : ((d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f))) * invn;
  }
  block_sum<1>(acc, sm);
  if (threadIdx.x == 0) out_loss[0] = acc[0] * invn;
}

// ref: finetune/ppo.py:865-874. Stable descending argsort by rank counting (n is the tag count: 2..80).
__global__ void ppo_rollout_kernel(const float* __restrict__ scores, const long long* __restrict__ state, int B, int n,
                                   int n_prefix, long long* __restrict__ next_state, long long* __restrict__ order) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* sb = scores + (long long)b * n;
  long long* ns = next_state + (long long)b * (n_prefix + n);
  for (int i = 0; i < n_prefix; ++i) ns[i] = i;
  for (int i = 0; i < n; ++i) {
    const float si = sb[i];
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      const float sj = sb[j];
      rank += (sj > si || (sj == si && j < i)) ? 1 : 0;
    }
    ns[n_prefix + rank] = state ? state[(long long)b * n + i] : (long long)i;
    if (order != nullptr) order[(long long)b * n + rank] = i;
  }
}

// Plackett-Luce sequential sampler with supplied uniforms; deterministic exp/log (det_math.h) so the
// permutation is bit-reproducible against the C oracle.
__global__ void rank_sample_kernel(const float* __restrict__ scores, const float* __restrict__ u, int B, int n,
                                   int greedy, long long* __restrict__ perm, float* __restrict__ logprob) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* sb = scores + (long long)b * n;
  long long* pb = perm + (long long)b * n;
  // perm doubles as the "taken" marker: initialise to -1, fill position by position.
  for (int t = 0; t < n; ++t) pb[t] = -1;
  float lp = 0.f;
  for (int t = 0; t < n; ++t) {
    float mx = -INFINITY;
    for (int j = 0; j < n; ++j) {
      bool taken = false;
      for (int q = 0; q < t; ++q) taken |= (pb[q] == j);
      if (!taken) mx = fmaxf(mx, sb[j]);
    }
    float total = 0.f;
    for (int j = 0; j < n; ++j) {
      bool taken = false;
      for (int q = 0; q < t; ++q) taken |= (pb[q] == j);
      if (!taken) total = __fadd_rn(total, lr2_det_expf(__fsub_rn(sb[j], mx)));
    }
    int pick = -1;
    float pick_e = 0.f;
    if (greedy) {
      for (int j = 0; j < n; ++j) {
        bool taken = false;
        for (int q = 0; q < t; ++q) taken |= (pb[q] == j);
        if (!taken && (pick < 0 || sb[j] > sb[pick])) pick = j;
      }
      pick_e = lr2_det_expf(__fsub_rn(sb[pick], mx));
    } else {
      const float target = __fmul_rn(u[(long long)b * n + t], total);
      float cum = 0.f;
      int last = -1;
      float last_e = 0.f;
      for (int j = 0; j < n; ++j) {
        bool taken = false;
        for (int q = 0; q < t; ++q) taken |= (pb[q] == j);
        if (taken) continue;
        const float e = lr2_det_expf(__fsub_rn(sb[j], mx));
        cum = __fadd_rn(cum, e);
        last = j; last_e = e;
        if (target < cum) { pick = j; pick_e = e; break; }
      }
      if (pick < 0) { pick = last; pick_e = last_e; }  // u*total rounded up to total
    }
    pb[t] = pick;
    lp = __fadd_rn(lp, __fsub_rn(lr2_det_logf(pick_e), lr2_det_logf(total)));
  }
  if (logprob != nullptr) logprob[b] = lp;
}

// Plackett-Luce log-probability of a GIVEN ranking under (new) scores, with the gradient: the quantity a PPO ratio
// needs.  Same arithmetic as rank_sample_kernel, step by step (max over the not-yet-placed labels, sum of det_exp in
// label-index order, det_log(e_pick) - det_log(total)), so rank_logprob(scores, rank_sample(scores).perm) reproduces the
// sampler's own logprob bit for bit and the ratio of an unchanged policy is exactly 1.
//   lp = sum_t [ (s[pi_t] - mx_t) - log sum_{j not placed before t} exp(s_j - mx_t) ]
//   d lp / d s_j = [j placed at k] - sum_{t <= k} exp(s_j - mx_t) / total_t          (k = position of j in pi)
// inv [B, n] (caller-owned scratch) receives the inverse permutation (position of each label).
__global__ void rank_logprob_kernel(const float* __restrict__ scores, const long long* __restrict__ perm, int B, int n,
                                    const float* __restrict__ dlp, float* __restrict__ logprob,
                                    float* __restrict__ dscores, int* __restrict__ inv) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* sb = scores + (long long)b * n;
  const long long* pb = perm + (long long)b * n;
  int* ib = inv + (long long)b * n;
  for (int t = 0; t < n; ++t) ib[(int)pb[t]] = t;
  float* db = dscores != nullptr ? dscores + (long long)b * n : nullptr;
  const float up = dlp != nullptr ? dlp[b] : 1.0f;
  if (db != nullptr)
    for (int j = 0; j < n; ++j) db[j] = up;            // the "+1" of every label (each is placed exactly once)
  float lp = 0.f;
  for (int t = 0; t < n; ++t) {
    float mx = -INFINITY;
    for (int j = 0; j < n; ++j)
      if (ib[j] >= t) mx = fmaxf(mx, sb[j]);
    float total = 0.f;
    for (int j = 0; j < n; ++j)
      if (ib[j] >= t) total = __fadd_rn(total, lr2_det_expf(__fsub_rn(sb[j], mx)));
    const float pick_e = lr2_det_expf(__fsub_rn(sb[(int)pb[t]], mx));
    lp = __fadd_rn(lp, __fsub_rn(lr2_det_logf(pick_e), lr2_det_logf(total)));
    if (db != nullptr) {
      const float inv_total = 1.0f / total;
      for (int j = 0; j < n; ++j)
        if (ib[j] >= t) db[j] -= up * (lr2_det_expf(__fsub_rn(sb[j], mx)) * inv_total);
    }
  }
  if (logprob != nullptr) logprob[b] = lp;
}

// Ratio-clipped PPO surrogate (north_star extension; the reference parses --eps_clip but never reads it, and keeps the
// PaLM-rlhf helper it would use, masked_normalize, as dead code: finetune/ppo.py:485-491):
//   A' = normalize ? (A - mean(A)) * rsqrt(max(mean((A - mean)^2), norm_eps)) : A                     (:485-491)
//   ratio = exp(logp - logp_old);  loss = -mean_b min(ratio * A', clamp(ratio, 1 - eps, 1 + eps) * A')
// out[0] = loss, out[1] = fraction of rows with |ratio - 1| > eps.  dlogp = d loss / d logp (advantages are constants).
// Ties of the two branches split the gradient evenly, as torch.min does (inside the clip range both are identical).
__global__ void __launch_bounds__(PL_THREADS)
ppo_clip_surrogate_kernel(const float* __restrict__ logp, const float* __restrict__ logp_old,
                          const float* __restrict__ adv, int B, float eps_clip, int normalize, float norm_eps,
                          float* __restrict__ out, float* __restrict__ dlogp, float* __restrict__ adv_used) {
  __shared__ float sm[2 * (PL_THREADS / 32)];
  const float invB = 1.f / (float)B;
  float mean = 0.f, scale = 1.f;
  if (normalize) {
    float a1[1] = {0.f};
    for (int b = threadIdx.x; b < B; b += PL_THREADS) a1[0] += adv[b];
    block_sum<1>(a1, sm);
    mean = a1[0] * invB;
    float a2[1] = {0.f};
    for (int b = threadIdx.x; b < B; b += PL_THREADS) { const float c = adv[b] - mean; a2[0] += c * c; }
    block_sum<1>(a2, sm);
    scale = 1.0f / sqrtf(fmaxf(a2[0] * invB, norm_eps));
  }
  float acc[2] = {0.f, 0.f};
  for (int b = threadIdx.x; b < B; b += PL_THREADS) {
    const float A = (adv[b] - mean) * scale;
    const float ratio = expf(logp[b] - logp_old[b]);
    const float lo = 1.f - eps_clip, hi = 1.f + eps_clip;
    const float rc = fminf(fmaxf(ratio, lo), hi);
    const float s1 = ratio * A, s2 = rc * A;
    acc[0] += fminf(s1, s2);
    acc[1] += (ratio < lo || ratio > hi) ? 1.f : 0.f;
    if (adv_used != nullptr) adv_used[b] = A;
    if (dlogp != nullptr) {
      const float g1 = A * ratio;                                        // d s1 / d logp
      const float g2 = (ratio >= lo && ratio <= hi) ? A * ratio : 0.f;   // d s2 / d logp (clamp passes inside only)
      float g;
      if (s1 < s2) g = g1;
      else if (s1 > s2) g = g2;
      else g = 0.5f * g1 + 0.5f * g2;
      dlogp[b] = -g * invB;
    }
  }
  block_sum<2>(acc, sm);
  if (threadIdx.x == 0) { out[0] = -acc[0] * invB; out[1] = acc[1] * invB; }
}

// GAE as a warp-shuffle reverse scan over affine maps A_t = d_t + c_t * A_{t+1}.
__global__ void gae_scan_kernel(const float* __restrict__ rewards, const float* __restrict__ values,
                                const float* __restrict__ notdone, int B, int T, float gamma, float lam,
                                float* __restrict__ adv, float* __restrict__ ret) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const float* r = rewards + (long long)b * T;
  const float* v = values + (long long)b * (T + 1);
  float carry = 0.f;  // A_{t+1} beyond the current chunk
  for (int hi = T; hi > 0; hi -= 32) {
    const int t = hi - 32 + lane;  // lanes cover [hi-32, hi)
    float a = 0.f, d = 0.f;        // identity for t < 0: A = 0*next + 0 (never stored)
    if (t >= 0) {
      const float nd = notdone ? notdone[(long long)b * T + t] : 1.f;
      d = r[t] + gamma * v[t + 1] * nd - v[t];
      a = gamma * lam * nd;
    }
    // inclusive reverse scan: combine (a,d) with the aggregate of lanes above
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float a2 = __shfl_down_sync(0xffffffffu, a, o);
      const float d2 = __shfl_down_sync(0xffffffffu, d, o);
      if (lane + o < 32) { d = d + a * d2; a = a * a2; }
    }
    const float A = d + a * carry;
    if (t >= 0) {
      adv[(long long)b * T + t] = A;
      if (ret != nullptr) ret[(long long)b * T + t] = A + v[t];
    }
    carry = __shfl_sync(0xffffffffu, A, 0);
    // lanes with t < 0 carry the identity composed with everything above: harmless, loop ends.
  }
}

}  // namespace lr2

using namespace lr2;
#define S_(x) reinterpret_cast<cudaStream_t>(x)

extern "C" int lr2_ppo_policy_loss(const float* s, const float* s_old, const float* reward, const float* v_old,
                                   const long long* pi, int B, int n, int k, float w_kl, float w_ent, float margin,
                                   float adv_eps, float* out_scalars, float* kl, float* ent, float* reward_adj,
                                   float* adv, float* ds, void* stream) {
  if (B <= 0 || n <= 0 || k <= 0) return LR2_ERR_BAD_SHAPE;
  ppo_policy_loss_kernel<<<1, PL_THREADS, 0, S_(stream)>>>(s, s_old, reward, v_old, pi, B, n, k, w_kl, w_ent, margin,
                                                           adv_eps, out_scalars, kl, ent, reward_adj, adv, ds); LR2_LAUNCHED(1);
  LR2_RETURN_LAUNCH();
}
extern "C" int lr2_clipped_value_loss(const float* v, const float* ret, const float* v_old, int B, float clip,
                                      float* out_loss, float* dv, void* stream) {
  if (B <= 0) return LR2_ERR_BAD_SHAPE;
  clipped_value_loss_kernel<<<1, PL_THREADS, 0, S_(stream)>>>(v, ret, v_old, B, clip, out_loss, dv); LR2_LAUNCHED(1);
  LR2_RETURN_LAUNCH();
}
extern "C" int lr2_pair_hinge_loss(const float* chosen, const float* reject, int B, float margin, float* out,
                                   float* dchosen, float* dreject, void* stream) {
  if (B <= 0) return LR2_ERR_BAD_SHAPE;
  pair_hinge_kernel<<<1, PL_THREADS, 0, S_(stream)>>>(chosen, reject, B, margin, out, dchosen, dreject); LR2_LAUNCHED(1);
  LR2_RETURN_LAUNCH();
}
extern "C" int lr2_smooth_l1_loss(const float* logits, const long long* tgt, long long n, float beta, float* out_loss,
                                  float* dlogits, void* stream) {
  if (n <= 0 || beta <= 0.f) return LR2_ERR_BAD_SHAPE;
  smooth_l1_kernel<<<1, PL_THREADS, 0, S_(stream)>>>(logits, tgt, n, beta, out_loss, dlogits); LR2_LAUNCHED(1);
  LR2_RETURN_LAUNCH();
}
extern "C" int lr2_ppo_rollout(const float* scores, const long long* state, int B, int n, int n_prefix,
                               long long* next_state, long long* order, void* stream) {
  if (B <= 0 || n <= 0 || n_prefix < 0) return LR2_ERR_BAD_SHAPE;
  ppo_rollout_kernel<<<(B + 127) / 128, 128, 0, S_(stream)>>>(scores, state, B, n, n_prefix, next_state, order); LR2_LAUNCHED(1);
  LR2_RETURN_LAUNCH();
}
extern "C" int lr2_rank_sample(const float* scores, const float* u, int B, int n, int greedy, long long* perm,
                               float* logprob, void* stream) {
  if (B <= 0 || n <= 0) return LR2_ERR_BAD_SHAPE;
  if (!greedy && u == nullptr) return LR2_ERR_BAD_SHAPE;
  rank_sample_kernel<<<(B + 127) / 128, 128, 0, S_(stream)>>>(scores, u, B, n, greedy, perm, logprob); LR2_LAUNCHED(1);
  LR2_RETURN_LAUNCH();
}
extern "C" int lr2_rank_logprob(const float* scores, const long long* perm, int B, int n, const float* dlogprob,
                                float* logprob, float* dscores, int* inv_scratch, void* stream) {
  if (B <= 0 || n <= 0) return LR2_ERR_BAD_SHAPE;
  if (scores == nullptr || perm == nullptr || inv_scratch == nullptr) return LR2_ERR_BAD_SHAPE;
  rank_logprob_kernel<<<(B + 127) / 128, 128, 0, S_(stream)>>>(scores, perm, B, n, dlogprob, logprob, dscores,
                                                               inv_scratch); LR2_LAUNCHED(1);
  LR2_RETURN_LAUNCH();
}
extern "C" int lr2_ppo_clip_surrogate(const float* logp, const float* logp_old, const float* adv, int B, float eps_clip,
                                      int normalize, float norm_eps, float* out, float* dlogp, float* adv_used,
                                      void* stream) {
  if (B <= 0 || eps_clip < 0.f) return LR2_ERR_BAD_SHAPE;
  if (logp == nullptr || logp_old == nullptr || adv == nullptr || out == nullptr) return LR2_ERR_BAD_SHAPE;
  ppo_clip_surrogate_kernel<<<1, PL_THREADS, 0, S_(stream)>>>(logp, logp_old, adv, B, eps_clip, normalize, norm_eps, out,
                                                              dlogp, adv_used); LR2_LAUNCHED(1);
  LR2_RETURN_LAUNCH();
}
extern "C" int lr2_gae_scan(const float* rewards, const float* values, const float* notdone, int B, int T,
                            float gamma, float lam, float* adv, float* ret, void* stream) {
  if (B <= 0 || T <= 0) return LR2_ERR_BAD_SHAPE;
  gae_scan_kernel<<<(B + 3) / 4, 128, 0, S_(stream)>>>(rewards, values, notdone, B, T, gamma, lam, adv, ret); LR2_LAUNCHED(1);
  LR2_RETURN_LAUNCH();
}
