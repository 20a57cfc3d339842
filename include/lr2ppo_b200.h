/*
 * lr2ppo_b200 — C ABI of the B200-native (sm_100a) kernels behind the LR2PPO
 * training / evaluation hot path.
 *
 * Conventions:
 *   - every pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - memory is caller-owned: kernels never allocate or free;
 *   - `stream` is a cudaStream_t passed as void*; all work is asynchronous on it;
 *   - return 0 on success, a negative LR2_ERR_* code otherwise; no exceptions, no exit;
 *   - "bf16" buffers hold __nv_bfloat16, "f32" float, "i64" long long;
 *   - citations `ref:` point into the reference tree (ChazzyGordon/LR2PPO) and name
 *     the Python code path the entry point replaces.
 *
 * The reference is pure Python/PyTorch: its "FFI" for this path is the set of
 * ATen calls inside the cited functions.  The binding a maintainer adds is the
 * ctypes layer in lr2ppo_b200/_lib.py.
 */
#ifndef LR2PPO_B200_H
#define LR2PPO_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define LR2_ABI_VERSION 1

/* error codes */
#define LR2_OK 0
#define LR2_ERR_BAD_SHAPE (-1)
#define LR2_ERR_BAD_DTYPE (-2)
#define LR2_ERR_MISALIGNED (-3)
#define LR2_ERR_WRONG_ARCH (-4)
#define LR2_ERR_CUDA (-5)
#define LR2_ERR_TMA (-6)
#define LR2_ERR_UNSUPPORTED (-7)

/* GEMM epilogues (output coordinates r = row, c = column of the written tensor) */
#define LR2_EPI_NONE 0          /* out = acc (+ beta*out_old for f32 outputs)                */
#define LR2_EPI_BIAS 1          /* out = acc + bias[c]                                        */
#define LR2_EPI_BIAS_GELU 2     /* pre = acc + bias[c]; C2 = pre; out = dropout(gelu_erf(pre)) */
#define LR2_EPI_BIAS_DROP_RES 3 /* out = dropout(acc + bias[c]) + aux[r,c]                    */
#define LR2_EPI_DGELU 4         /* out = acc * gelu_erf'(aux[r,c]) * dropout_mask             */
#define LR2_EPI_ADD 5           /* out = acc + aux[r,c]                                       */
#define LR2_EPI_ADAMW 6         /* internal: fused wgrad + AdamW (lr2_gemm_wgrad_adamw)       */
#define LR2_EPI_BIAS_QGELU 7    /* like BIAS_GELU with QuickGELU x*sigmoid(1.702x) (video_transformer.py:91) */
#define LR2_EPI_DQGELU 8        /* like DGELU with QuickGELU'                                 */

int lr2_abi_version(void);
const char* lr2_last_error_string(int code);
/* 0 when the current device is compute capability 10.x, LR2_ERR_WRONG_ARCH otherwise. */
int lr2_check_device(void);
/* number of CUDA kernels this library has launched in the current process */
long long lr2_launch_count(void);
void lr2_note_launches(int n);

/* ---------------------------------------------------------------- dense --
 * D[M,N] = A[M,K] * B[N,K]^T, bf16 x bf16 -> fp32 (tcgen05.mma, TMEM accumulators, TMA loads).
 * a_mn_major / b_mn_major = 0: operand stored row-major [rows, K] (pitch lda/ldb elements);
 *                         = 1: operand stored row-major [K, rows].
 * transposed_out = 1 writes element (m, n) at C[n*ldc + m] (epilogue tensors follow the
 * written orientation).  splits > 1 = split-K through `workspace`
 * (lr2_gemm_workspace_bytes).  block_n in {0 (auto), 64, 128, 256} selects the single-CTA kernel's N tile;
 * 2128 / 2256 select the cta_group::2 pair kernel (a two-CTA cluster computes 256 x 128 / 256 x 256 tiles).  Auto uses the
 * pair kernel for untransposed problems with N % 256 == 0, K > 128 and at least 37 pair tiles (x splits).
 * Plain bf16 outputs of the 256-wide pair kernel (no epilogue, no split-K) leave through a TMA store
 * (cp.async.bulk.tensor global <- shared, {64, 32} boxes clipped at the matrix edges); every other case through
 * the staged register epilogue.
 * Dropout (all epilogues, lr2_layernorm_bwd, lr2_dropout_bf16): ONE Philox4x32-7 call per aligned group of 8 output
 * elements (linear index r*ldc + c), 16 random bits per element; p is quantised to round(p * 65536) / 65536 and the
 * survivors are scaled by exactly 65536 / (65536 - round(p * 65536)), so the mask is a pure function of
 * (seed [+ *seed_dev], site, element index) and is regenerated wherever backward needs it.
 * ref: every nn.Linear on the path — finetune/ppo.py:154-170 (Mlp), :207-208 (out_layer),
 *      finetune/xit.py:103-147 (FeedForwardBlock, MultiHeadAttention projections),
 *      tencentpretrain/layers/multi_headed_attn.py:27-76, position_ffn.py:12-15 — and their
 *      autograd backward (dgrad / wgrad).
 */
long long lr2_gemm_workspace_bytes(int M, int N, int splits, int transposed_out, long long ldc);
int lr2_gemm_bf16(const void* A, long long lda, int a_mn_major, const void* B, long long ldb, int b_mn_major,
                  void* C, long long ldc, int c_is_f32, int transposed_out, int M, int N, int K, int epilogue,
                  const float* bias, const void* aux, long long ldaux, void* C2, float beta, float drop_p,
                  unsigned long long seed, unsigned int site, const void* seed_dev, int splits, void* workspace,
                  int block_n,
                  void* stream);

/* Fused weight-gradient + AdamW for one Linear weight [out_f, in_f] (used for out_layer.fc1, 500 M params):
 * grad = dY[rows, out_f]^T @ X[rows, in_f] stays in TMEM and is consumed by the AdamW update of
 * param / exp_avg / exp_avg_sq (fp32, in place) and the bf16 shadow; no gradient tensor exists.
 * hyper = the 8-float device buffer of lr2_adamw_multi.
 * ref: finetune/ppo.py:579-580 (loss.backward(); optimizer.step()) restricted to out_layer.fc1.weight;
 *      tencentpretrain/utils/optimizers.py:374-402. */
int lr2_gemm_wgrad_adamw(const void* dY, long long lddy, const void* X, long long ldx, int rows, int out_f, int in_f,
                         float* param, float* exp_avg, float* exp_avg_sq, void* shadow_bf16, const float* hyper,
                         float weight_decay, void* stream);

/* ------------------------------------------------------------ layernorm --
 * mode 0: torch nn.LayerNorm (biased variance, eps inside sqrt)   ref: finetune/xit.py:31-41,71-74,96-100
 * mode 1: TencentPretrain LayerNorm gamma*(x-mean)/(std_unbiased+eps)+beta
 *                                                                  ref: tencentpretrain/layers/layer_norm.py:16-21
 * Rows of y may be re-grouped: out_row = (row / g_in) * g_out + row % g_in + g_off (g_in <= 0: identity);
 * this is how the final XiT LayerNorm writes straight into the [items, 196+16, 768] concat buffer
 * (ref: finetune/ppo.py:224 torch.cat).  stats[2*row] = mean, stats[2*row+1] = 1/sigma.
 */
int lr2_layernorm_fwd(const void* x_bf16, const float* gamma, const float* beta, void* y_bf16, float* stats,
                      long long rows, int D, float eps, int mode, int g_in, int g_out, int g_off, void* stream);
/* dx = LN'(dy) (+ add_bf16 if non-null).  dy rows use the same re-grouping as y in fwd.
 * dxm (optional) = dx * dropout_mask(seed, site)/(1-p): the gradient that flows into the
 * dropout-ed branch feeding this LayerNorm's input.  dgamma/dbeta are accumulated through
 * `partials` (>= lr2_layernorm_bwd_partials_floats(D) floats). */
long long lr2_layernorm_bwd_partials_floats(int D);
int lr2_layernorm_bwd(const void* dy_bf16, const void* x_bf16, const float* gamma, const float* stats,
                      const void* add_bf16, void* dx_bf16, void* dxm_bf16, float* dgamma, float* dbeta,
                      float* partials, long long rows, int D, float eps, int mode, int g_in, int g_out, int g_off,
                      float drop_p, unsigned long long seed, unsigned int site, const void* seed_dev, void* stream);

/* -------------------------------------------------- XiT attention core --
 * Per (item, head): P = softmax(pre_scale * Q K^T); O = (post_scale * P) V, Skv <= 16.
 * XiT uses pre_scale = 1, post_scale = 1/sqrt(emb)  (softmax-then-scale quirk, no mask)
 *                                                                  ref: finetune/xit.py:125-148
 * q: [items, Sq, H*dh] pitch ldq; k, v: [items, Skv, H*dh] pitch ldkv; o pitch ldo. dh % 8 == 0, dh <= 128.
 */
int lr2_xattn_fwd(const void* q, long long ldq, const void* k, const void* v, long long ldkv, void* o,
                  long long ldo, int items, int Sq, int Skv, int H, int dh, float pre_scale, float post_scale,
                  void* stream);
int lr2_xattn_bwd(const void* q, long long ldq, const void* k, const void* v, long long ldkv, const void* d_o,
                  long long ldo, void* dq, long long lddq, void* dk, void* dv, long long lddkv, int items, int Sq,
                  int Skv, int H, int dh, float pre_scale, float post_scale, void* stream);

/* ------------------------------------- TencentPretrain tower kernels --
 * Multi-headed self-attention core, flash-style (scores never leave the SM), S <= 256, dh == 64:
 *   P = softmax(Q K^T * scale + key_bias[b, j]);  O = dropout(P) V;  lse[b,h,i] saved for backward.
 * q/k/v: [B*S, ...] rows with pitch ld (views into a merged QKV buffer are fine), head h at column h*dh.
 * ref: tencentpretrain/layers/multi_headed_attn.py:55-76, encoders/transformer_encoder.py:62-68 (mask). */
int lr2_mha_fwd(const void* q, const void* k, const void* v, long long ld, const float* key_bias, void* o,
                long long ldo, float* lse, int B, int S, int H, int dh, float scale, float drop_p,
                unsigned long long seed, const void* seed_dev, void* stream);
int lr2_mha_bwd(const void* q, const void* k, const void* v, long long ld, const float* key_bias, const void* o,
                const void* d_o, long long ldo, const float* lse, void* dq, void* dk, void* dv, long long ldd, int B,
                int S, int H, int dh, float scale, float drop_p, unsigned long long seed, const void* seed_dev,
                void* stream);
/* out[r,:] = word[src[r],:] + pos[r % S,:] (+ seg_table[seg[r],:])   ref: embeddings/{word,pos,seg}_embedding.py */
int lr2_embed_sum(const long long* src, const long long* seg, const float* word, const float* pos,
                  const float* seg_table, void* out_bf16, long long rows, int S, int D, void* stream);
/* table[idx[r],:] += d[r,:] (fp32 atomics): gradient of an embedding lookup */
int lr2_embed_scatter_add(const long long* idx, const void* d_bf16, float* table, long long rows, int D, void* stream);
/* im2col for Conv2d(k = s = ps, no bias): [B,C,H,W] f32 -> [B*(H/ps)*(W/ps), C*ps*ps] bf16
 * ref: embeddings/patch_embedding.py:18,27 */
int lr2_patchify(const float* img, void* out_bf16, int B, int C, int Hh, int Ww, int ps, void* stream);
/* out = gelu(bf16(x + bias)), pre (optional) = bf16(x + bias): the bias + exact-GELU epilogue of Mlp.fc1
 * (finetune/ppo.py:164-166) applied to fp32 pre-activation sums that were reduced over the ranks (K-split
 * out_layer.fc1, lr2ppo_b200/dist.py); same arithmetic as the LR2_EPI_BIAS_GELU epilogue.  x fp32 [rows, D]. */
int lr2_bias_gelu_rows(const float* x, const float* bias, void* out_bf16, void* pre_bf16, long long rows, int D,
                       void* stream);
/* out = x * keep(seed, site, element index) / (1-p): elementwise dropout / its backward (n % 8 == 0) */
int lr2_dropout_bf16(const void* x, void* out, long long n, float p, unsigned long long seed, unsigned int site,
                     const void* seed_dev, void* stream);

/* ------------------------------------------------------ glue kernels --- */
/* counter[0] += inc (u64, device): the dropout seed offset read through `seed_dev` above; bumping it inside a
 * captured CUDA graph gives every replay fresh masks (the reference draws new nn.Dropout masks per call). */
int lr2_bump_counter(void* counter, unsigned long long inc, void* stream);
/* dst[b, j, :] = bf16(src[b, index[b, j], :]); index == NULL -> identity (T_dst == T_src).
 * ref: finetune/ppo.py:268-271 (text_emb[batch_index, index]) fused with the fp32 -> bf16 cast. */
int lr2_cast_gather_bf16(const float* src, const long long* index, void* dst_bf16, int bs, int T_src, int T_dst,
                         long long row_elems, void* stream);
/* dst[b, j, :] = src[b, index[b, j], :] on bf16 rows (feature reuse when index repeats items: the reward
 * model's 4-slot input [0, 1, pi(0), pi(1)] holds only 2 distinct items, ref: finetune/ppo.py:318-322). */
int lr2_gather_rows_bf16(const void* src, const long long* index, void* dst, int bs, int T_src, int T_dst,
                         long long row_elems, void* stream);
/* grouped row copy: dst[(g*dst_gstride + dst_off + r), :] (+)= src[(g*src_gstride + src_off + r), :]
 * ref: finetune/ppo.py:224 (cat of x and img_feature) and its backward split. */
int lr2_rows_copy_bf16(const void* src, long long src_gstride, long long src_off, void* dst, long long dst_gstride,
                       long long dst_off, long long groups, long long rows_per_group, int D, int accumulate,
                       void* stream);
/* out[c] (+)= sum_r x[r, c]  (bias gradients).  partials >= lr2_colsum_partials_floats(cols) floats. */
long long lr2_colsum_partials_floats(int cols);
int lr2_colsum_bf16(const void* x, long long ldx, long long rows, int cols, float* out, float* partials,
                    int accumulate, void* stream);
/* out[r] = dot(x[r*row_stride + row_off, :], w) + b[0]       ref: finetune/ppo.py:228,293-295 (head, last token) */
int lr2_rowdot_fwd(const void* x_bf16, long long row_stride, long long row_off, const float* w, const float* b,
                   float* out, int rows, int D, void* stream);
/* dx (all rows*row_stride rows, zero where not selected), dw[D], db[1] */
int lr2_rowdot_bwd(const void* x_bf16, long long row_stride, long long row_off, const float* w, const float* dout,
                   void* dx_bf16, float* dw, float* db, int rows, int D, void* stream);
/* x[b, t, :] += pos[t, :]                                    ref: finetune/ppo.py:286-289 */
int lr2_add_pos_fwd(void* x_bf16, const float* pos, int bs, int T, int D, void* stream);
int lr2_add_pos_bwd(const void* dx_bf16, float* dpos, int bs, int T, int D, void* stream);
int lr2_cast_f32_to_bf16(const float* src, void* dst, long long n, void* stream);
int lr2_cast_bf16_to_f32(const void* src, float* dst, long long n, void* stream);

/* ------------------------------------------------------------- PPO rows --
 * Stage-3 update losses, forward + analytic backward in one launch.
 * ref: finetune/ppo.py:38-55 (RankLoss), :431-432 (log), :544-575 (KL, entropy, advantage, policy loss).
 * s, s_old: [B, n] f32; reward, v_old: [B]; pi: [B, k] i64 ranked index list with values in [0, n)
 * (the update passes next_state[:, -2:], k = 2; RankLoss accepts any k, finetune/ppo.py:43-46).
 * out_scalars[0..3] = {policy_loss, rank_loss, hinge_cnt, sum|adv|}; per-row outputs [B] each;
 * ds = d policy_loss / d s  [B, n].  An index outside [0, n) is never dereferenced: policy_loss and rank_loss
 * come back NaN.
 */
int lr2_ppo_policy_loss(const float* s, const float* s_old, const float* reward, const float* v_old,
                        const long long* pi, int B, int n, int k, float w_kl, float w_ent, float margin, float adv_eps,
                        float* out_scalars, float* kl, float* ent, float* reward_adj, float* adv, float* ds,
                        void* stream);
/* ref: finetune/ppo.py:494-498.  out_loss[0] = mean(max((vc-R)^2,(V-R)^2)); dv = d loss / d V. */
int lr2_clipped_value_loss(const float* v, const float* ret, const float* v_old, int B, float clip, float* out_loss,
                           float* dv, void* stream);
/* ref: finetune/reward_pair_dataloader.py:355-358 (margin 1), finetune/reward_trad.py:273 (margin 0.01).
 * out[0] = mean(relu(margin-(c-r))), out[1] = mean(c > r). */
int lr2_pair_hinge_loss(const float* chosen, const float* reject, int B, float margin, float* out, float* dchosen,
                        float* dreject, void* stream);
/* ref: finetune/pointwise.py:229 nn.SmoothL1Loss(beta=0.3) on (logits, int64 targets). */
int lr2_smooth_l1_loss(const float* logits, const long long* tgt, long long n, float beta, float* out_loss,
                       float* dlogits, void* stream);
/* Rollout: idx = stable argsort_desc(scores); next_state = [0..n_prefix-1, state[idx]].
 * ref: finetune/ppo.py:865-874.  state may be NULL (= arange(n)).  next_state: [B, n_prefix+n] i64. */
int lr2_ppo_rollout(const float* scores, const long long* state, int B, int n, int n_prefix, long long* next_state,
                    long long* order, void* stream);
/* Sequential masked-softmax (Plackett-Luce) ranking sampler with caller-supplied uniforms u[B, n] in [0,1).
 * greedy != 0 ignores u and takes the arg-max at every position (== lr2_ppo_rollout order).
 * Outputs perm [B, n] i64 and logprob [B] f32.  north_star extension; oracle: oracle/ppo_rows.c. */
int lr2_rank_sample(const float* scores, const float* u, int B, int n, int greedy, long long* perm, float* logprob,
                    void* stream);
/* Plackett-Luce log-probability of a given ranking perm [B,n] under scores [B,n] (north_star extension, the
 * counterpart of lr2_rank_sample: same arithmetic step by step, so evaluating the sampler's own ranking under the same
 * scores reproduces its logprob bit for bit).  logprob [B] and/or dscores [B,n] = dlogprob[b] * d lp_b / d scores
 * (dlogprob NULL = 1).  inv_scratch: caller-owned int [B,n].  No reference implementation exists: the reference ranks
 * greedily (finetune/ppo.py:865-874); log / masked helpers it keeps as dead code: finetune/ppo.py:431-491. */
int lr2_rank_logprob(const float* scores, const long long* perm, int B, int n, const float* dlogprob, float* logprob,
                     float* dscores, int* inv_scratch, void* stream);

/* Ratio-clipped PPO surrogate, forward + gradient in one launch (north_star extension; --eps_clip is parsed but never
 * read by the reference, finetune/ppo.py:730, and masked_normalize is dead code at finetune/ppo.py:485-491):
 * A' = normalize ? (A - mean A) * rsqrt(max(var A, norm_eps)) : A;  ratio = exp(logp - logp_old);
 * out[0] = -mean min(ratio A', clamp(ratio, 1 - eps, 1 + eps) A'), out[1] = clipped fraction;
 * dlogp [B] (or NULL), adv_used [B] (or NULL) = A'. */
int lr2_ppo_clip_surrogate(const float* logp, const float* logp_old, const float* adv, int B, float eps_clip,
                           int normalize, float norm_eps, float* out, float* dlogp, float* adv_used, void* stream);

/* GAE(gamma, lambda) reverse scan: delta_t = r_t + gamma*V_{t+1}*nd_t - V_t; A_t = delta_t + gamma*lambda*nd_t*A_{t+1}.
 * rewards [B,T], values [B,T+1], notdone [B,T] or NULL; adv, ret [B,T].  T=1, V_1=0 -> r - V (ref: finetune/ppo.py:560). */
int lr2_gae_scan(const float* rewards, const float* values, const float* notdone, int B, int T, float gamma,
                 float lam, float* adv, float* ret, void* stream);

/* ----------------------------------------------------------------- NDCG --
 * Per query q (row pitch ld, length len[q] or N): sort scores descending (stable), gather labels,
 * ideal = labels sorted descending, DCG@k accumulated sequentially in fp32 exactly like
 * ref: ndcg.py:28-32,54-65 with callers finetune/ppo.py:651-659.
 * log2_table[i] = fp32 log2(i+2).  ks [nk] i64.  ndcg [B, nk] f32.  order [B, ld] i64 (optional).
 * N <= 4096, nk <= 32.  gain = float(int64(2**label - 1)); labels outside [0, 63] give 2**label == 0 (gain -1).
 * N <= 1024 runs one warp per query (register/shuffle bitonic network) unless few long queries are given;
 * LR2_NDCG_LEGACY=1 forces the block-per-query kernel, LR2_NDCG_WARP=1 the warp kernel (both bit-identical).
 * LR2_NDCG_BLOCK_PAIRS=1 (opt-in) lets the block kernel index compare-exchange pairs directly (no idle threads).
 */
int lr2_ndcg_at_k(const float* scores, const long long* labels, const int* lens, int B, int N, long long ld,
                  const long long* ks, int nk, const float* log2_table, float* ndcg, long long* order,
                  void* stream);

/* Same arithmetic on two relevance lists that are ALREADY in rank order — the exact signature of
 * AverageNDCGMeter.return_ndcg_at_k(predicted_relevance, true_relevances) (ref: ndcg.py:54-65).
 * pred_rel, true_rel: [B, N] i64; scratch: 2*B*nk floats. */
int lr2_ndcg_presorted(const long long* pred_rel, const long long* true_rel, const int* lens, int B, int N,
                       const long long* ks, int nk, const float* log2_table, float* ndcg, float* scratch,
                       void* stream);

/* ---------------------------------------------------------------- AdamW --
 * HF-style AdamW without bias correction, decay applied after the Adam update with lr*wd.
 * ref: tencentpretrain/utils/optimizers.py:344-402.
 * The tensor table lives in device memory: for tensor t, ptrs[6*t+0..5] =
 *   {p f32, g, m f32, v f32, shadow bf16 or NULL, unused}; meta[4*t+0..3] = {n, wd (float bits), g_is_bf16, 0}.
 * chunks[2*c+0..1] = {tensor id | (length << 32), element offset}; length 0 = the default chunk of
 * lr2_adamw_chunk_elems() elements (or up to the tensor's end); an explicit length (multiple of 4, at most the default)
 * is a piece of a column block of a 2-D parameter (column-sharded data-parallel optimizer).
 * hyper (device, 8 floats) = {step_size, beta1, beta2, eps, 1-beta1, 1-beta2, grad_scale, lr_for_decay}
 * (step_size == lr when correct_bias is False, as in every LR2PPO script).
 */
int lr2_adamw_chunk_elems(void);
int lr2_adamw_multi(const void* const* ptrs, const long long* meta, const long long* chunks, long long num_chunks,
                    const float* hyper, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LR2PPO_B200_H */
