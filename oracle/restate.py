"""TEST INFRASTRUCTURE ONLY — fp32 torch/numpy restatements of the reference's row losses, optimizer
and normalisation (SURVEY.md §8a').  Each function cites the reference lines it follows.  Validated
against the imported reference by oracle/make_golden.py -> tests/golden/rows.json.
"""
import ctypes
import math
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
C_LIB = os.path.join(_HERE, "_build", "liboracle.so")


def build_c_oracle():
    """gcc the C restatement (rows.c). Building the checker is not using it."""
    import subprocess
    os.makedirs(os.path.dirname(C_LIB), exist_ok=True)
    src = os.path.join(_HERE, "rows.c")
    if os.path.exists(C_LIB) and os.path.getmtime(C_LIB) >= os.path.getmtime(src):
        return C_LIB
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", src, "-o", C_LIB,
                           "-lm"])
    return C_LIB


_c = None


def c_oracle():
    global _c
    if _c is None:
        _c = ctypes.CDLL(build_c_oracle())
    return _c


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def log2_table(n):
    """fp32 log2(i+2) through the same torch call as ndcg.py:31 (`torch.log2(i + 2)`, i int64)."""
    return torch.log2(torch.arange(2, n + 2, dtype=torch.int64)).numpy().astype(np.float32)


# ----------------------------------------------------------------- NDCG ---
def ndcg_at_k(scores, labels, ks, lens=None, want_order=False):
    """ref: ndcg.py:28-32,54-65 + finetune/ppo.py:651-659.  numpy in / numpy out, via rows.c."""
    scores = np.ascontiguousarray(scores, dtype=np.float32)
    labels = np.ascontiguousarray(labels, dtype=np.int64)
    B, N = scores.shape
    ks_a = np.asarray(list(ks), dtype=np.int64)
    out = np.empty((B, len(ks_a)), dtype=np.float32)
    order = np.empty((B, N), dtype=np.int64) if want_order else None
    lens_a = None if lens is None else np.ascontiguousarray(lens, dtype=np.int32)
    tab = log2_table(N)
    c_oracle().oracle_ndcg_at_k(_p(scores), _p(labels), _p(lens_a), B, N, _p(ks_a), len(ks_a), _p(tab), _p(out),
                                _p(order))
    return (out, order) if want_order else out


def ndcg_python(pred_rel, true_rel, ks):
    """Pure-python/numpy-fp32 transliteration of the arithmetic for SMALL cases (cross-check of rows.c)."""
    tab = log2_table(max(len(pred_rel), 1))

    def dcg(rel, k):
        acc = np.float32(0.0)
        for i in range(min(len(rel), k)):
            gain = np.float32(np.int64(2) ** np.int64(rel[i]) - np.int64(1))
            acc = np.float32(acc + np.float32(gain / tab[i]))
        return acc

    out = []
    for k in ks:
        p, t = dcg(pred_rel, k), dcg(true_rel, k)
        out.append(np.float32(1.0) if t <= np.float32(1e-6) else np.float32(p / t))
    return np.asarray(out, dtype=np.float32)


def ppo_rollout(scores, state=None, n_prefix=2):
    """ref: finetune/ppo.py:865-874."""
    scores = np.ascontiguousarray(scores, dtype=np.float32)
    B, n = scores.shape
    st = None if state is None else np.ascontiguousarray(state, dtype=np.int64)
    ns = np.empty((B, n_prefix + n), dtype=np.int64)
    order = np.empty((B, n), dtype=np.int64)
    c_oracle().oracle_ppo_rollout(_p(scores), _p(st), B, n, n_prefix, _p(ns), _p(order))
    return ns, order


def rank_sample(scores, u=None, greedy=False):
    scores = np.ascontiguousarray(scores, dtype=np.float32)
    B, n = scores.shape
    ua = None if u is None else np.ascontiguousarray(u, dtype=np.float32)
    perm = np.empty((B, n), dtype=np.int64)
    lp = np.empty(B, dtype=np.float32)
    c_oracle().oracle_rank_sample(_p(scores), _p(ua), B, n, int(greedy), _p(perm), _p(lp))
    return perm, lp


def gae(rewards, values, gamma, lam, notdone=None):
    rewards = np.ascontiguousarray(rewards, dtype=np.float32)
    values = np.ascontiguousarray(values, dtype=np.float32)
    B, T = rewards.shape
    nd = None if notdone is None else np.ascontiguousarray(notdone, dtype=np.float32)
    adv = np.empty((B, T), dtype=np.float32)
    ret = np.empty((B, T), dtype=np.float32)
    c_oracle().oracle_gae(_p(rewards), _p(values), _p(nd), B, T, ctypes.c_float(gamma), ctypes.c_float(lam), _p(adv),
                          _p(ret))
    return adv, ret


# --------------------------------------------------------------- losses ---
def clamp_log(t, eps=1e-20):
    """ref: finetune/ppo.py:431-432."""
    return torch.log(t.clamp(min=eps))


def rank_loss(scores, order, margin):
    """ref: finetune/ppo.py:43-55 (RankLoss.forward)."""
    s = torch.gather(scores, 1, order)
    diff = margin - (s.unsqueeze(2) - s.unsqueeze(1))
    hinge = torch.relu(torch.triu(diff, diagonal=1))
    cnt = torch.sign(hinge).sum()
    if cnt == 0:
        return hinge.sum(), cnt
    return hinge.sum() / cnt, cnt


def ppo_policy_loss(s, s_old, reward, v_old, pi, w_kl, w_ent, margin=0.01, adv_eps=-0.1):
    """ref: finetune/ppo.py:544-575.  s requires_grad for gradients.  Returns dict of tensors."""
    p_old = s_old.softmax(dim=-1)
    p = s.softmax(dim=-1)
    kl = (p_old * (clamp_log(p_old) - clamp_log(p))).sum(dim=-1)
    ent = -(p * clamp_log(p)).sum(dim=-1)
    radj = reward - kl * w_kl
    adv = radj - v_old
    rows = []
    for i in range(adv.shape[0]):
        rows.append(pi[i] if adv[i] >= adv_eps else pi[i].flip(dims=[-1]))
    order = torch.stack(rows)
    rl, cnt = rank_loss(s, order, margin)
    loss = (rl * adv.abs() - w_ent * ent).mean()
    return dict(loss=loss, rank_loss=rl, hinge_cnt=cnt, kl=kl, entropy=ent, reward_adj=radj, adv=adv)


def clipped_value_loss(values, rewards, old_values, clip):
    """ref: finetune/ppo.py:494-498."""
    vc = old_values + (values - old_values).clamp(-clip, clip)
    l1 = (vc.flatten() - rewards) ** 2
    l2 = (values.flatten() - rewards) ** 2
    return torch.mean(torch.max(l1, l2))


def pair_hinge_loss(chosen, reject, margin=1.0):
    """ref: finetune/reward_pair_dataloader.py:355-358 (margin 1), finetune/reward_trad.py:273 (0.01)."""
    loss = torch.relu(margin - (chosen - reject)).mean()
    acc = (chosen > reject).float().mean()
    return loss, acc


def smooth_l1(logits, tgt, beta=0.3):
    """ref: finetune/pointwise.py:229."""
    d = logits.view(-1) - tgt.view(-1).to(logits.dtype)
    ad = d.abs()
    return torch.where(ad < beta, 0.5 * d * d / beta, ad - 0.5 * beta).mean()


# ------------------------------------------------------------ optimizer ---
def adamw_step(p, g, m, v, lr, wd, beta1=0.9, beta2=0.999, eps=1e-6):
    """ref: tencentpretrain/utils/optimizers.py:374-402 (correct_bias=False). In-place on p, m, v."""
    m.mul_(beta1).add_(g, alpha=1.0 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1.0 - beta2)
    denom = v.sqrt().add_(eps)
    p.addcdiv_(m, denom, value=-lr)
    if wd > 0.0:
        p.add_(p, alpha=-lr * wd)


def linear_schedule_lambda(step, warmup_steps, total_steps):
    """ref: tencentpretrain/utils/optimizers.py:79-84."""
    if step < warmup_steps:
        return float(step) / float(max(1, warmup_steps))
    return max(0.0, float(total_steps - step) / float(max(1, total_steps - warmup_steps)))


# -------------------------------------------------------- normalisation ---
def tencent_layernorm(x, gamma, beta, eps=1e-6):
    """ref: tencentpretrain/layers/layer_norm.py:16-21."""
    mean = x.mean(-1, keepdim=True)
    std = x.std(-1, keepdim=True)
    return gamma * (x - mean) / (std + eps) + beta


# ---- north_star extensions: Plackett-Luce log-probability of a ranking and the ratio-clipped surrogate ------------
def masked_normalize(t, eps=1e-5):
    """ref: finetune/ppo.py:485-491 (dead code in the reference; mask / dim are ignored there too)."""
    mean = t.mean()
    c = t - mean
    var = (c ** 2).mean()
    return c * var.clamp(min=eps).rsqrt()


def rank_logprob(scores, perm):
    """log P(perm | scores) under the sequential masked-softmax (Plackett-Luce) model of oracle_rank_sample, as a
    differentiable torch expression (fp32): sum_t [ s[pi_t] - logsumexp_{j >= t} s[pi_j] ]."""
    s = torch.gather(scores, 1, perm)                       # scores in ranking order
    n = s.shape[1]
    lp = torch.zeros(s.shape[0], dtype=s.dtype)
    for t in range(n):
        lp = lp + s[:, t] - torch.logsumexp(s[:, t:], dim=1)
    return lp


def ppo_clip_surrogate(logp, logp_old, adv, eps_clip=0.2, normalize=False, norm_eps=1e-5):
    """PaLM-rlhf-style clipped surrogate built from the reference's helpers (log: finetune/ppo.py:431-432,
    masked_normalize: :485-491); the reference itself never reads --eps_clip (finetune/ppo.py:730).
    Returns (loss, clipped fraction); differentiable in logp."""
    a = masked_normalize(adv, norm_eps) if normalize else adv
    ratio = (logp - logp_old).exp()
    s1 = ratio * a
    s2 = ratio.clamp(1 - eps_clip, 1 + eps_clip) * a
    loss = -torch.min(s1, s2).mean()
    frac = ((ratio < 1 - eps_clip) | (ratio > 1 + eps_clip)).float().mean()
    return loss, frac
