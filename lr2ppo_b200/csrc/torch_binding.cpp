// torch custom-op layer over the C ABI (include/lr2ppo_b200.h): TORCH_LIBRARY(lr2ppo, m).
//
// "one C symbol + one torch.ops.lr2ppo.* each".  Every op here validates its tensors (CUDA, dtype,
// contiguity), allocates its outputs with torch, passes raw pointers and at::cuda::getCurrentCUDAStream() to the
// matching lr2_* entry point and turns a non-zero return code into a c10::Error (TORCH_CHECK).  Nothing is computed
// here and there is no fallback: a CPU tensor is an error.  The ops are CUDA-graph capturable (no synchronisation, no
// host-side reads of device memory).  Built by `make torch` into lr2ppo_b200/liblr2ppo_torch.so, which links
// liblr2ppo_b200.so (rpath $ORIGIN); lr2ppo_b200/torch_ops.py loads it with torch.ops.load_library.
//
// Names follow SURVEY.md §8(b)'s op list; the reference lines each op replaces are cited in the header next to its
// C symbol.
#include <ATen/ATen.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/library.h>

#include "../../include/lr2ppo_b200.h"

namespace {

using at::Tensor;
using OptT = const std::optional<Tensor>&;

void* stream() { return at::cuda::getCurrentCUDAStream().stream(); }

void ck(int rc, const char* what) { TORCH_CHECK(rc == 0, "lr2ppo::", what, ": error ", rc, " (", lr2_last_error_string(rc), ")"); }

const Tensor& cuda(const Tensor& t, at::ScalarType dt, const char* name) {
  TORCH_CHECK(t.is_cuda(), "lr2ppo: ", name, " must be a CUDA tensor (no CPU fallback)");
  TORCH_CHECK(t.scalar_type() == dt, "lr2ppo: ", name, " must be ", dt, ", got ", t.scalar_type());
  TORCH_CHECK(t.is_contiguous(), "lr2ppo: ", name, " must be contiguous");
  return t;
}
const void* optp(OptT t, at::ScalarType dt, const char* name) { return t.has_value() ? cuda(*t, dt, name).data_ptr() : nullptr; }
const auto BF = at::kBFloat16;
const auto F32 = at::kFloat;
const auto I64 = at::kLong;

// ---- gemm_bias_act: D[M,N] = A[M,K] B[N,K]^T with the fused epilogues (Linear fwd / dgrad / wgrad) -----------------
Tensor gemm(const Tensor& a, const Tensor& b, bool a_mn, bool b_mn, int64_t epilogue, OptT bias, OptT aux, OptT c2,
            bool out_f32, bool transposed_out, double drop_p, int64_t seed, int64_t site, int64_t splits,
            int64_t block_n) {
  TORCH_CHECK(a.is_cuda() && b.is_cuda() && a.scalar_type() == BF && b.scalar_type() == BF && a.dim() == 2 && b.dim() == 2 &&
                  a.stride(1) == 1 && b.stride(1) == 1, "lr2ppo::gemm: operands must be 2-D bf16 CUDA tensors with unit inner stride");
  c10::cuda::CUDAGuard guard(a.device());
  const int64_t M = a_mn ? a.size(1) : a.size(0), K = a_mn ? a.size(0) : a.size(1);
  const int64_t N = b_mn ? b.size(1) : b.size(0);
  TORCH_CHECK((b_mn ? b.size(0) : b.size(1)) == K, "lr2ppo::gemm: K mismatch");
  const int64_t rows = transposed_out ? N : M, cols = transposed_out ? M : N;
  Tensor out = at::empty({rows, cols}, a.options().dtype(out_f32 ? F32 : BF));
  Tensor ws;
  if (splits > 1) {
    ws = at::empty({lr2_gemm_workspace_bytes((int)M, (int)N, (int)splits, transposed_out, cols)}, a.options().dtype(at::kByte));
  }
  if (aux.has_value()) TORCH_CHECK(aux->scalar_type() == BF && aux->stride(-1) == 1, "lr2ppo::gemm: aux must be bf16");
  if (c2.has_value()) TORCH_CHECK(c2->scalar_type() == BF && c2->stride(-2) == cols, "lr2ppo::gemm: c2 must match the output pitch");
  ck(lr2_gemm_bf16(a.data_ptr(), a.stride(0), a_mn, b.data_ptr(), b.stride(0), b_mn, out.data_ptr(), cols, out_f32,
                   transposed_out, (int)M, (int)N, (int)K, (int)epilogue, (const float*)optp(bias, F32, "bias"),
                   aux.has_value() ? aux->data_ptr() : nullptr, aux.has_value() ? aux->stride(-2) : 0,
                   c2.has_value() ? c2->data_ptr() : nullptr, 0.f, (float)drop_p, (unsigned long long)seed,
                   (unsigned int)site, nullptr, (int)splits, ws.defined() ? ws.data_ptr() : nullptr, (int)block_n,
                   stream()), "gemm");
  return out;
}

// ---- layernorm (mode 0: nn.LayerNorm, 1: TencentPretrain std-based LayerNorm) ---------------------------------------
std::tuple<Tensor, Tensor> layernorm_fwd(const Tensor& x, const Tensor& gamma, const Tensor& beta, double eps, int64_t mode) {
  cuda(x, BF, "x"); cuda(gamma, F32, "gamma"); cuda(beta, F32, "beta");
  c10::cuda::CUDAGuard guard(x.device());
  const int64_t D = x.size(-1), rows = x.numel() / D;
  Tensor y = at::empty_like(x), stats = at::empty({rows, 2}, x.options().dtype(F32));
  ck(lr2_layernorm_fwd(x.data_ptr(), gamma.data_ptr<float>(), beta.data_ptr<float>(), y.data_ptr(), stats.data_ptr<float>(),
                       rows, (int)D, (float)eps, (int)mode, 0, 0, 0, stream()), "layernorm_fwd");
  return {y, stats};
}
std::tuple<Tensor, Tensor, Tensor> layernorm_bwd(const Tensor& dy, const Tensor& x, const Tensor& gamma,
                                                 const Tensor& stats, double eps, int64_t mode) {
  cuda(dy, BF, "dy"); cuda(x, BF, "x"); cuda(gamma, F32, "gamma"); cuda(stats, F32, "stats");
  c10::cuda::CUDAGuard guard(x.device());
  const int64_t D = x.size(-1), rows = x.numel() / D;
  Tensor dx = at::empty_like(x), dg = at::empty({D}, x.options().dtype(F32)), db = at::empty({D}, x.options().dtype(F32));
  Tensor part = at::empty({lr2_layernorm_bwd_partials_floats((int)D)}, x.options().dtype(F32));
  ck(lr2_layernorm_bwd(dy.data_ptr(), x.data_ptr(), gamma.data_ptr<float>(), stats.data_ptr<float>(), nullptr,
                       dx.data_ptr(), nullptr, dg.data_ptr<float>(), db.data_ptr<float>(), part.data_ptr<float>(), rows,
                       (int)D, (float)eps, (int)mode, 0, 0, 0, 0.f, 0ull, 0u, nullptr, stream()), "layernorm_bwd");
  return {dx, dg, db};
}

// ---- xit_attention: softmax(Q K^T * pre) * post  V  over tiny key sets (finetune/xit.py:133-146) -----------------
Tensor xit_attention_fwd(const Tensor& q, const Tensor& k, const Tensor& v, int64_t heads, double pre, double post) {
  cuda(q, BF, "q"); cuda(k, BF, "k"); cuda(v, BF, "v");
  c10::cuda::CUDAGuard guard(q.device());
  const int64_t items = q.size(0), Sq = q.size(1), E = q.size(2), Skv = k.size(1);
  Tensor o = at::empty_like(q);
  ck(lr2_xattn_fwd(q.data_ptr(), E, k.data_ptr(), v.data_ptr(), E, o.data_ptr(), E, (int)items, (int)Sq, (int)Skv,
                   (int)heads, (int)(E / heads), (float)pre, (float)post, stream()), "xit_attention_fwd");
  return o;
}
std::tuple<Tensor, Tensor, Tensor> xit_attention_bwd(const Tensor& q, const Tensor& k, const Tensor& v, const Tensor& d_o,
                                                     int64_t heads, double pre, double post) {
  cuda(q, BF, "q"); cuda(k, BF, "k"); cuda(v, BF, "v"); cuda(d_o, BF, "d_o");
  c10::cuda::CUDAGuard guard(q.device());
  const int64_t items = q.size(0), Sq = q.size(1), E = q.size(2), Skv = k.size(1);
  Tensor dq = at::empty_like(q), dk = at::empty_like(k), dv = at::empty_like(v);
  ck(lr2_xattn_bwd(q.data_ptr(), E, k.data_ptr(), v.data_ptr(), E, d_o.data_ptr(), E, dq.data_ptr(), E, dk.data_ptr(),
                   dv.data_ptr(), E, (int)items, (int)Sq, (int)Skv, (int)heads, (int)(E / heads), (float)pre, (float)post,
                   stream()), "xit_attention_bwd");
  return {dq, dk, dv};
}

// ---- flash_attention of the TencentPretrain towers: qkv [B*S, 3E] merged projections ---------------------------------
std::tuple<Tensor, Tensor> flash_attention_fwd(const Tensor& qkv, int64_t B, int64_t S, int64_t H, OptT key_bias, double scale,
                                               double drop_p, int64_t seed) {
  cuda(qkv, BF, "qkv");
  c10::cuda::CUDAGuard guard(qkv.device());
  const int64_t E = qkv.size(1) / 3;
  Tensor o = at::empty({B * S, E}, qkv.options()), lse = at::empty({B, H, S}, qkv.options().dtype(F32));
  char* base = (char*)qkv.data_ptr();
  ck(lr2_mha_fwd(base, base + 2 * E, base + 4 * E, qkv.stride(0), (const float*)optp(key_bias, F32, "key_bias"),
                 o.data_ptr(), E, lse.data_ptr<float>(), (int)B, (int)S, (int)H, (int)(E / H), (float)scale, (float)drop_p,
                 (unsigned long long)seed, nullptr, stream()), "flash_attention_fwd");
  return {o, lse};
}
Tensor flash_attention_bwd(const Tensor& qkv, const Tensor& o, const Tensor& d_o, const Tensor& lse, int64_t B, int64_t S,
                           int64_t H, OptT key_bias, double scale, double drop_p, int64_t seed) {
  cuda(qkv, BF, "qkv"); cuda(o, BF, "o"); cuda(d_o, BF, "d_o"); cuda(lse, F32, "lse");
  c10::cuda::CUDAGuard guard(qkv.device());
  const int64_t E = qkv.size(1) / 3;
  Tensor d = at::empty_like(qkv);
  char* base = (char*)qkv.data_ptr();
  char* db = (char*)d.data_ptr();
  ck(lr2_mha_bwd(base, base + 2 * E, base + 4 * E, qkv.stride(0), (const float*)optp(key_bias, F32, "key_bias"),
                 o.data_ptr(), d_o.data_ptr(), E, lse.data_ptr<float>(), db, db + 2 * E, db + 4 * E, d.stride(0), (int)B,
                 (int)S, (int)H, (int)(E / H), (float)scale, (float)drop_p, (unsigned long long)seed, nullptr, stream()),
     "flash_attention_bwd");
  return d;
}

// ---- gather_items: text_emb[batch_index, index] fused with the fp32 -> bf16 cast (finetune/ppo.py:268-271) ---------
Tensor gather_items(const Tensor& src, OptT index) {
  cuda(src, F32, "src");
  c10::cuda::CUDAGuard guard(src.device());
  const int64_t bs = src.size(0), Tsrc = src.size(1), row = src.numel() / (bs * Tsrc);
  const int64_t Tdst = index.has_value() ? cuda(*index, I64, "index").size(1) : Tsrc;
  auto shape = src.sizes().vec();
  shape[1] = Tdst;
  Tensor out = at::empty(shape, src.options().dtype(BF));
  ck(lr2_cast_gather_bf16(src.data_ptr<float>(), index.has_value() ? (const long long*)index->data_ptr<int64_t>() : nullptr, out.data_ptr(),
                          (int)bs, (int)Tsrc, (int)Tdst, row, stream()), "gather_items");
  return out;
}

Tensor bias_gelu(const Tensor& x, const Tensor& bias) {
  cuda(x, F32, "x"); cuda(bias, F32, "bias");
  c10::cuda::CUDAGuard guard(x.device());
  Tensor out = at::empty(x.sizes(), x.options().dtype(BF));
  ck(lr2_bias_gelu_rows(x.data_ptr<float>(), bias.data_ptr<float>(), out.data_ptr(), nullptr, x.size(0), (int)x.size(1),
                        stream()), "bias_gelu");
  return out;
}

Tensor dropout_philox(const Tensor& x, double p, int64_t seed, int64_t site) {
  cuda(x, BF, "x");
  c10::cuda::CUDAGuard guard(x.device());
  Tensor out = at::empty_like(x);
  ck(lr2_dropout_bf16(x.data_ptr(), out.data_ptr(), x.numel(), (float)p, (unsigned long long)seed, (unsigned int)site,
                      nullptr, stream()), "dropout_philox");
  return out;
}

// ---- PPO / ranking row kernels -----------------------------------------------------------------------------------
Tensor ppo_rollout(const Tensor& scores, const Tensor& state, int64_t n_prefix) {
  cuda(scores, F32, "scores"); cuda(state, I64, "state");
  c10::cuda::CUDAGuard guard(scores.device());
  const int64_t B = scores.size(0), n = scores.size(1);
  Tensor ns = at::empty({B, n_prefix + n}, scores.options().dtype(I64));
  ck(lr2_ppo_rollout(scores.data_ptr<float>(), (const long long*)state.data_ptr<int64_t>(), (int)B, (int)n, (int)n_prefix,
                     (long long*)ns.data_ptr<int64_t>(), nullptr, stream()), "ppo_rollout");
  return ns;
}
// -> (scalars[4] = {policy_loss, rank_loss, hinge_cnt, sum|adv|}, kl[B], entropy[B], reward_adj[B], adv[B], ds[B,n])
std::tuple<Tensor, Tensor, Tensor, Tensor, Tensor, Tensor> ppo_policy_loss(const Tensor& s, const Tensor& s_old,
                                                                          const Tensor& reward, const Tensor& v_old,
                                                                          const Tensor& pi, double w_kl, double w_ent,
                                                                          double margin, double adv_eps) {
  cuda(s, F32, "s"); cuda(s_old, F32, "s_old"); cuda(reward, F32, "reward"); cuda(v_old, F32, "v_old"); cuda(pi, I64, "pi");
  TORCH_CHECK(s.dim() == 2 && s_old.sizes() == s.sizes() && pi.dim() == 2 && pi.size(0) == s.size(0) &&
                  reward.numel() == s.size(0) && v_old.numel() == s.size(0), "lr2ppo::ppo_policy_loss: shape mismatch");
  c10::cuda::CUDAGuard guard(s.device());
  const int64_t B = s.size(0), n = s.size(1);
  auto f = s.options();
  Tensor scal = at::empty({4}, f), kl = at::empty({B}, f), ent = at::empty({B}, f), radj = at::empty({B}, f),
         adv = at::empty({B}, f), ds = at::empty({B, n}, f);
  ck(lr2_ppo_policy_loss(s.data_ptr<float>(), s_old.data_ptr<float>(), reward.data_ptr<float>(), v_old.data_ptr<float>(),
                         (const long long*)pi.data_ptr<int64_t>(), (int)B, (int)n, (int)pi.size(1), (float)w_kl, (float)w_ent,
                         (float)margin, (float)adv_eps, scal.data_ptr<float>(), kl.data_ptr<float>(), ent.data_ptr<float>(),
                         radj.data_ptr<float>(), adv.data_ptr<float>(), ds.data_ptr<float>(), stream()), "ppo_policy_loss");
  return {scal, kl, ent, radj, adv, ds};
}
std::tuple<Tensor, Tensor> clipped_value_loss(const Tensor& v, const Tensor& ret, const Tensor& v_old, double clip) {
  cuda(v, F32, "v"); cuda(ret, F32, "ret"); cuda(v_old, F32, "v_old");
  c10::cuda::CUDAGuard guard(v.device());
  Tensor loss = at::empty({1}, v.options()), dv = at::empty({v.numel()}, v.options());
  ck(lr2_clipped_value_loss(v.data_ptr<float>(), ret.data_ptr<float>(), v_old.data_ptr<float>(), (int)v.numel(), (float)clip,
                            loss.data_ptr<float>(), dv.data_ptr<float>(), stream()), "clipped_value_loss");
  return {loss, dv};
}
// -> (out[2] = {loss, accuracy}, dchosen[B], dreject[B])
std::tuple<Tensor, Tensor, Tensor> pair_hinge_loss(const Tensor& chosen, const Tensor& reject, double margin) {
  cuda(chosen, F32, "chosen"); cuda(reject, F32, "reject");
  c10::cuda::CUDAGuard guard(chosen.device());
  const int64_t B = chosen.numel();
  Tensor out = at::empty({2}, chosen.options()), dc = at::empty({B}, chosen.options()), dr = at::empty({B}, chosen.options());
  ck(lr2_pair_hinge_loss(chosen.data_ptr<float>(), reject.data_ptr<float>(), (int)B, (float)margin, out.data_ptr<float>(),
                         dc.data_ptr<float>(), dr.data_ptr<float>(), stream()), "pair_hinge_loss");
  return {out, dc, dr};
}
std::tuple<Tensor, Tensor> smooth_l1(const Tensor& logits, const Tensor& tgt, double beta) {
  cuda(logits, F32, "logits"); cuda(tgt, I64, "tgt");
  c10::cuda::CUDAGuard guard(logits.device());
  Tensor loss = at::empty({1}, logits.options()), dl = at::empty({logits.numel()}, logits.options());
  ck(lr2_smooth_l1_loss(logits.data_ptr<float>(), (const long long*)tgt.data_ptr<int64_t>(), logits.numel(), (float)beta,
                        loss.data_ptr<float>(), dl.data_ptr<float>(), stream()), "smooth_l1");
  return {loss, dl};
}
std::tuple<Tensor, Tensor> gae_scan(const Tensor& rewards, const Tensor& values, double gamma, double lam) {
  cuda(rewards, F32, "rewards"); cuda(values, F32, "values");
  TORCH_CHECK(values.size(0) == rewards.size(0) && values.size(1) == rewards.size(1) + 1, "lr2ppo::gae_scan: values must be [B, T+1]");
  c10::cuda::CUDAGuard guard(rewards.device());
  Tensor adv = at::empty_like(rewards), ret = at::empty_like(rewards);
  ck(lr2_gae_scan(rewards.data_ptr<float>(), values.data_ptr<float>(), nullptr, (int)rewards.size(0), (int)rewards.size(1),
                  (float)gamma, (float)lam, adv.data_ptr<float>(), ret.data_ptr<float>(), stream()), "gae_scan");
  return {adv, ret};
}

// ---- ndcg_at_k: segmented sort + sequential fp32 DCG (ndcg.py:28-65); log2_table = torch.log2(arange(2, N + 2)) -----
std::tuple<Tensor, Tensor> ndcg_at_k(const Tensor& scores, const Tensor& labels, const Tensor& ks, const Tensor& log2_table,
                                     OptT lens) {
  cuda(scores, F32, "scores"); cuda(labels, I64, "labels"); cuda(ks, I64, "ks"); cuda(log2_table, F32, "log2_table");
  c10::cuda::CUDAGuard guard(scores.device());
  const int64_t B = scores.size(0), N = scores.size(1), nk = ks.numel();
  Tensor out = at::empty({B, nk}, scores.options()), order = at::full({B, N}, -1, scores.options().dtype(I64));
  ck(lr2_ndcg_at_k(scores.data_ptr<float>(), (const long long*)labels.data_ptr<int64_t>(),
                   lens.has_value() ? cuda(*lens, at::kInt, "lens").data_ptr<int>() : nullptr, (int)B, (int)N, N,
                   (const long long*)ks.data_ptr<int64_t>(), (int)nk, log2_table.data_ptr<float>(), out.data_ptr<float>(),
                   (long long*)order.data_ptr<int64_t>(), stream()), "ndcg_at_k");
  return {out, order};
}

// ---- adamw_multi_tensor: one launch over a device-resident tensor table (see the header for the table layout) -------
void adamw_multi_tensor(const Tensor& ptrs, const Tensor& meta, const Tensor& chunks, const Tensor& hyper) {
  cuda(ptrs, I64, "ptrs"); cuda(meta, I64, "meta"); cuda(chunks, I64, "chunks"); cuda(hyper, F32, "hyper");
  c10::cuda::CUDAGuard guard(ptrs.device());
  ck(lr2_adamw_multi((const void* const*)ptrs.data_ptr<int64_t>(), (const long long*)meta.data_ptr<int64_t>(),
                     (const long long*)chunks.data_ptr<int64_t>(), chunks.size(0), hyper.data_ptr<float>(), stream()),
     "adamw_multi_tensor");
}

}  // namespace

TORCH_LIBRARY(lr2ppo, m) {
  m.def("gemm(Tensor a, Tensor b, bool a_mn=False, bool b_mn=False, int epilogue=0, Tensor? bias=None, Tensor? aux=None, "
        "Tensor? c2=None, bool out_f32=False, bool transposed_out=False, float drop_p=0.0, int seed=0, int site=0, "
        "int splits=1, int block_n=0) -> Tensor");
  m.def("layernorm_fwd(Tensor x, Tensor gamma, Tensor beta, float eps, int mode=0) -> (Tensor, Tensor)");
  m.def("layernorm_bwd(Tensor dy, Tensor x, Tensor gamma, Tensor stats, float eps, int mode=0) -> (Tensor, Tensor, Tensor)");
  m.def("xit_attention_fwd(Tensor q, Tensor k, Tensor v, int heads, float pre_scale, float post_scale) -> Tensor");
  m.def("xit_attention_bwd(Tensor q, Tensor k, Tensor v, Tensor d_o, int heads, float pre_scale, float post_scale) -> "
        "(Tensor, Tensor, Tensor)");
  m.def("flash_attention_fwd(Tensor qkv, int B, int S, int H, Tensor? key_bias, float scale, float drop_p=0.0, int seed=0) -> "
        "(Tensor, Tensor)");
  m.def("flash_attention_bwd(Tensor qkv, Tensor o, Tensor d_o, Tensor lse, int B, int S, int H, Tensor? key_bias, float scale, "
        "float drop_p=0.0, int seed=0) -> Tensor");
  m.def("gather_items(Tensor src, Tensor? index=None) -> Tensor");
  m.def("bias_gelu(Tensor x, Tensor bias) -> Tensor");
  m.def("dropout_philox(Tensor x, float p, int seed, int site) -> Tensor");
  m.def("ppo_rollout(Tensor scores, Tensor state, int n_prefix=2) -> Tensor");
  m.def("ppo_policy_loss(Tensor s, Tensor s_old, Tensor reward, Tensor v_old, Tensor pi, float w_kl, float w_ent, "
        "float margin=0.01, float adv_eps=-0.1) -> (Tensor, Tensor, Tensor, Tensor, Tensor, Tensor)");
  m.def("clipped_value_loss(Tensor v, Tensor ret, Tensor v_old, float clip) -> (Tensor, Tensor)");
  m.def("pair_hinge_loss(Tensor chosen, Tensor reject, float margin=1.0) -> (Tensor, Tensor, Tensor)");
  m.def("smooth_l1(Tensor logits, Tensor tgt, float beta=0.3) -> (Tensor, Tensor)");
  m.def("gae_scan(Tensor rewards, Tensor values, float gamma, float lam) -> (Tensor, Tensor)");
  m.def("ndcg_at_k(Tensor scores, Tensor labels, Tensor ks, Tensor log2_table, Tensor? lens=None) -> (Tensor, Tensor)");
  m.def("adamw_multi_tensor(Tensor ptrs, Tensor meta, Tensor chunks, Tensor hyper) -> ()");
}

TORCH_LIBRARY_IMPL(lr2ppo, CUDA, m) {
  m.impl("gemm", &gemm);
  m.impl("layernorm_fwd", &layernorm_fwd);
  m.impl("layernorm_bwd", &layernorm_bwd);
  m.impl("xit_attention_fwd", &xit_attention_fwd);
  m.impl("xit_attention_bwd", &xit_attention_bwd);
  m.impl("flash_attention_fwd", &flash_attention_fwd);
  m.impl("flash_attention_bwd", &flash_attention_bwd);
  m.impl("gather_items", &gather_items);
  m.impl("bias_gelu", &bias_gelu);
  m.impl("dropout_philox", &dropout_philox);
  m.impl("ppo_rollout", &ppo_rollout);
  m.impl("ppo_policy_loss", &ppo_policy_loss);
  m.impl("clipped_value_loss", &clipped_value_loss);
  m.impl("pair_hinge_loss", &pair_hinge_loss);
  m.impl("smooth_l1", &smooth_l1);
  m.impl("gae_scan", &gae_scan);
  m.impl("ndcg_at_k", &ndcg_at_k);
  m.impl("adamw_multi_tensor", &adamw_multi_tensor);
}
