/* TEST INFRASTRUCTURE — CPU restatement (oracle) of the integer / sequential-fp32 row algorithms
 * on the LR2PPO hot path.  Never linked into the product.
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC rows.c -o _build/liboracle.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "../lr2ppo_b200/csrc/det_math.h"

/* ---- stable descending argsort (ties: lower index first), insertion sort ---- */
static void argsort_desc_f32(const float* s, int n, int64_t* idx) {
  for (int i = 0; i < n; ++i) idx[i] = i;
  for (int i = 1; i < n; ++i) {
    int64_t v = idx[i];
    int j = i - 1;
    while (j >= 0 && s[idx[j]] < s[v]) { idx[j + 1] = idx[j]; --j; }
    idx[j + 1] = v;
  }
}

/* ref: ndcg.py:28-32 — dcg = 0; for i < min(len,k): dcg += (2**rel[i] - 1) / log2(i+2)
 * (2**rel - 1) is int64, log2 is fp32, the quotient and the running sum are fp32. */
static float gain_f32(int64_t rel) {
  int64_t g;
  if (rel < 0 || rel >= 64) g = -1;
  else g = (int64_t)((UINT64_C(1) << rel) - UINT64_C(1));
  return (float)g;
}
static float dcg_at_k(const int64_t* rel, int n, int64_t k, const float* log2_table) {
  float dcg = 0.0f;
  int64_t cut = k < n ? k : n;
  for (int64_t i = 0; i < cut; ++i) {
    volatile float term = gain_f32(rel[i]) / log2_table[i];
    dcg = dcg + term;
  }
  return dcg;
}

/* ref: ndcg.py:54-65 + callers finetune/ppo.py:651-659.
 * scores [B,N] f32, labels [B,N] i64, lens [B] (or NULL), ks [nk]; out ndcg [B,nk], order [B,N] (or NULL). */
void oracle_ndcg_at_k(const float* scores, const int64_t* labels, const int32_t* lens, int B, int N,
                      const int64_t* ks, int nk, const float* log2_table, float* ndcg, int64_t* order) {
  int64_t* idx = (int64_t*)malloc(sizeof(int64_t) * N);
  int64_t* pred = (int64_t*)malloc(sizeof(int64_t) * N);
  int64_t* ideal = (int64_t*)malloc(sizeof(int64_t) * N);
  for (int q = 0; q < B; ++q) {
    int n = lens ? (lens[q] < N ? lens[q] : N) : N;
    const float* s = scores + (size_t)q * N;
    const int64_t* l = labels + (size_t)q * N;
    argsort_desc_f32(s, n, idx);
    for (int i = 0; i < n; ++i) pred[i] = l[idx[i]];
    if (order) {
      for (int i = 0; i < N; ++i) order[(size_t)q * N + i] = i < n ? idx[i] : -1;
    }
    memcpy(ideal, l, sizeof(int64_t) * n);
    for (int i = 1; i < n; ++i) { /* descending insertion sort */
      int64_t v = ideal[i];
      int j = i - 1;
      while (j >= 0 && ideal[j] < v) { ideal[j + 1] = ideal[j]; --j; }
      ideal[j + 1] = v;
    }
    for (int j = 0; j < nk; ++j) {
      float p = dcg_at_k(pred, n, ks[j], log2_table);
      float t = dcg_at_k(ideal, n, ks[j], log2_table);
      ndcg[(size_t)q * nk + j] = (t <= 1e-6f) ? 1.0f : p / t;
    }
  }
  free(idx); free(pred); free(ideal);
}

/* ref: finetune/ppo.py:865-874 — idx = argsort_desc(scores); next_state = [0..n_prefix-1, state[idx]] */
void oracle_ppo_rollout(const float* scores, const int64_t* state, int B, int n, int n_prefix, int64_t* next_state,
                        int64_t* order) {
  int64_t* idx = (int64_t*)malloc(sizeof(int64_t) * n);
  for (int b = 0; b < B; ++b) {
    argsort_desc_f32(scores + (size_t)b * n, n, idx);
    int64_t* ns = next_state + (size_t)b * (n_prefix + n);
    for (int i = 0; i < n_prefix; ++i) ns[i] = i;
    for (int i = 0; i < n; ++i) {
      ns[n_prefix + i] = state ? state[(size_t)b * n + idx[i]] : idx[i];
      if (order) order[(size_t)b * n + i] = idx[i];
    }
  }
  free(idx);
}

/* north_star extension (no reference code): sequential masked-softmax (Plackett-Luce) ranking sampler
 * with caller-supplied uniforms.  Position t draws from softmax over the not-yet-placed labels by
 * inverse CDF: pick the first remaining j (index order) with u*total < cumsum_j.
 * greedy: take arg-max (first on ties) -> equals the stable descending sort of the reference. */
void oracle_rank_sample(const float* scores, const float* u, int B, int n, int greedy, int64_t* perm,
                        float* logprob) {
  char* taken = (char*)malloc(n);
  for (int b = 0; b < B; ++b) {
    const float* s = scores + (size_t)b * n;
    memset(taken, 0, n);
    float lp = 0.0f;
    for (int t = 0; t < n; ++t) {
      float mx = -INFINITY;
      for (int j = 0; j < n; ++j) if (!taken[j] && s[j] > mx) mx = s[j];
      float total = 0.0f;
      for (int j = 0; j < n; ++j) if (!taken[j]) total = total + lr2_det_expf(s[j] - mx);
      int pick = -1;
      float pick_e = 0.0f;
      if (greedy) {
        for (int j = 0; j < n; ++j) if (!taken[j] && (pick < 0 || s[j] > s[pick])) pick = j;
        pick_e = lr2_det_expf(s[pick] - mx);
      } else {
        float target = u[(size_t)b * n + t] * total;
        float cum = 0.0f, last_e = 0.0f;
        int last = -1;
        for (int j = 0; j < n; ++j) {
          if (taken[j]) continue;
          float e = lr2_det_expf(s[j] - mx);
          cum = cum + e;
          last = j; last_e = e;
          if (target < cum) { pick = j; pick_e = e; break; }
        }
        if (pick < 0) { pick = last; pick_e = last_e; }
      }
      taken[pick] = 1;
      perm[(size_t)b * n + t] = pick;
      lp = lp + (lr2_det_logf(pick_e) - lr2_det_logf(total));
    }
    if (logprob) logprob[b] = lp;
  }
  free(taken);
}

/* north_star extension: GAE(gamma, lambda), sequential reverse recursion.
 * delta_t = r_t + gamma*V_{t+1}*nd_t - V_t ; A_t = delta_t + gamma*lambda*nd_t*A_{t+1}
 * T = 1 with V_1 = 0 reduces to r - V (ref: finetune/ppo.py:560). */
void oracle_gae(const float* rewards, const float* values, const float* notdone, int B, int T, float gamma, float lam,
                float* adv, float* ret) {
  for (int b = 0; b < B; ++b) {
    float next = 0.0f;
    for (int t = T - 1; t >= 0; --t) {
      float nd = notdone ? notdone[(size_t)b * T + t] : 1.0f;
      float v = values[(size_t)b * (T + 1) + t], v1 = values[(size_t)b * (T + 1) + t + 1];
      float delta = rewards[(size_t)b * T + t] + gamma * v1 * nd - v;
      next = delta + gamma * lam * nd * next;
      adv[(size_t)b * T + t] = next;
      if (ret) ret[(size_t)b * T + t] = next + v;
    }
  }
}
