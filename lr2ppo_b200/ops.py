"""Tensor-level wrappers over the C ABI (one function per entry point).

Each wrapper validates device/dtype/contiguity, allocates outputs with torch, passes raw
pointers and the current CUDA stream to liblr2ppo_b200.so and raises on a non-zero return.
No wrapper computes anything on the host.
"""
import math
import torch

from . import _lib
from ._lib import ptr, check

EPI_NONE, EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_DROP_RES, EPI_DGELU, EPI_ADD = range(6)
EPI_BIAS_QGELU, EPI_DQGELU = 7, 8   # QuickGELU variants (6 is the internal fused-AdamW epilogue)

bf16 = torch.bfloat16
f32 = torch.float32
i64 = torch.int64


def _L():
    return _lib.load()


def _cuda(t, dtype=None, name="tensor"):
    if t is None:
        return
    if not t.is_cuda:
        raise _lib.Lr2Error(f"{name} must be a CUDA tensor (no CPU fallback)")
    if dtype is not None and t.dtype != dtype:
        raise _lib.Lr2Error(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise _lib.Lr2Error(f"{name} must be contiguous")


_ws_cache = {}


def _workspace(nbytes, device):
    """Grow-only scratch buffer per device (split-K slabs, reduction partials)."""
    key = (device.index, torch.cuda.current_stream().cuda_stream)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


# ------------------------------------------------------------------ GEMM --
def plan_gemm(M, N, K, transposed_out=False):
    """Split-K factor for an untransposed [M,N] problem, matching the kernel selection in lr2_gemm_bf16: problems whose
    256 x 256 pair tiles would leave most of the 74 CTA pairs idle (weight gradients: few output tiles, long K) are
    split along K until the pairs are filled; otherwise 1."""
    if transposed_out or N % 256 or M < 256 or K <= 128:
        return 1
    t = ((M + 255) // 256) * (N // 256)
    if t >= 74:
        return 1
    s = max(1, min(74 // t, ((K + 63) // 64) // 8, 16))
    return s if t * s >= 37 else 1


def plan_small_gemm(M, N, K):
    """Split-K factor for problems with a handful of output tiles and a long reduction (out_layer.fc2 on <= 96 rows,
    the xitt projections, K/V projections): a few CTAs walking K serially are DRAM-latency bound (25 us for
    48x768x3072 on 6 CTAs), so K is spread over up to 148 CTAs and folded by lr2_splitk_reduce."""
    tiles = ((M + 127) // 128) * ((N + 127) // 128)
    kb = (K + 63) // 64
    if tiles > 36 or kb < 16 or N % 8:
        return 1
    return max(1, min(16, 148 // tiles, kb // 4))


def gemm(a, b, *, a_mn=False, b_mn=False, out=None, out_dtype=bf16, transposed_out=False, epilogue=EPI_NONE,
         bias=None, aux=None, c2=None, beta=0.0, drop_p=0.0, seed=0, site=0, seed_dev=None, splits=None, block_n=0,
         M=None, N=None, K=None):
    """D[M,N] = A[M,K] @ B[N,K]^T  (bf16 in, fp32 accumulate).

    a: [M,K] (a_mn=False) or [K,M] (a_mn=True); b likewise with N.  Row pitch = stride(0).
    Output is [M,N], or [N,M] when transposed_out.
    """
    L = _L()
    for t, nm in ((a, "a"), (b, "b")):
        if not t.is_cuda or t.dtype != bf16 or t.dim() != 2 or t.stride(1) != 1:
            raise _lib.Lr2Error(f"gemm operand {nm} must be a 2-D bf16 CUDA tensor with unit inner stride")
    if M is None:
        M = a.shape[1] if a_mn else a.shape[0]
    if K is None:
        K = a.shape[0] if a_mn else a.shape[1]
    if N is None:
        N = b.shape[1] if b_mn else b.shape[0]
    kb = b.shape[0] if b_mn else b.shape[1]
    if kb != K:
        raise _lib.Lr2Error(f"gemm K mismatch: {K} vs {kb}")
    rows, cols = (N, M) if transposed_out else (M, N)
    if out is None:
        out = torch.empty((rows, cols), dtype=out_dtype, device=a.device)
    if out.shape[-2:] != (rows, cols) or out.stride(-1) != 1:
        raise _lib.Lr2Error("gemm: bad output shape")
    ldc = out.stride(-2) if out.dim() >= 2 else cols
    if splits is None:      # auto: fill the CTA pairs / SMs when the output has few tiles (same rule as the C side)
        splits = 1
        if not transposed_out and block_n == 0 and beta == 0.0:
            splits = plan_gemm(M, N, K)
            if splits == 1:
                splits = plan_small_gemm(M, N, K)
    ws = None
    if splits > 1:
        nbytes = L.lr2_gemm_workspace_bytes(M, N, splits, int(transposed_out), ldc)
        ws = _workspace(nbytes, a.device)
    _cuda(bias, f32, "bias")
    if aux is not None and (aux.dtype != bf16 or aux.stride(-1) != 1):
        raise _lib.Lr2Error("aux must be bf16 with unit inner stride")
    if c2 is not None and (c2.dtype != bf16 or c2.stride(-2) != ldc):
        raise _lib.Lr2Error("c2 must be bf16 with the same pitch as the output")
    _lib.run(L.lr2_gemm_bf16, ptr(a), a.stride(0), int(a_mn), ptr(b), b.stride(0), int(b_mn), ptr(out), ldc,
                          int(out.dtype == f32), int(transposed_out), M, N, K, epilogue, ptr(bias), ptr(aux),
                          aux.stride(-2) if aux is not None else 0, ptr(c2), float(beta), float(drop_p), int(seed),
                          int(site), ptr(seed_dev), int(splits), ptr(ws), int(block_n), _lib.stream())
    return out


# ------------------------------------------------------------- layernorm --
def layernorm_fwd(x, gamma, beta, eps, mode=0, out=None, regroup=None, want_stats=True):
    L = _L()
    _cuda(x, bf16, "x"); _cuda(gamma, f32, "gamma"); _cuda(beta, f32, "beta")
    D = x.shape[-1]
    rows = x.numel() // D
    g_in, g_out, g_off = regroup if regroup else (0, 0, 0)
    if out is None:
        out = torch.empty_like(x)
    stats = torch.empty((rows, 2), dtype=f32, device=x.device) if want_stats else None
    _lib.run(L.lr2_layernorm_fwd, ptr(x), ptr(gamma), ptr(beta), ptr(out), ptr(stats), rows, D, float(eps), mode, g_in,
                              g_out, g_off, _lib.stream())
    return out, stats


def layernorm_bwd(dy, x, gamma, stats, eps, mode=0, add=None, regroup=None, drop_p=0.0, seed=0, site=0,
                  want_masked=False, seed_dev=None):
    L = _L()
    _cuda(dy, bf16, "dy"); _cuda(x, bf16, "x"); _cuda(gamma, f32, "gamma"); _cuda(stats, f32, "stats")
    _cuda(add, bf16, "add")
    D = x.shape[-1]
    rows = x.numel() // D
    g_in, g_out, g_off = regroup if regroup else (0, 0, 0)
    dx = torch.empty_like(x)
    dxm = torch.empty_like(x) if want_masked else None
    dgamma = torch.empty(D, dtype=f32, device=x.device)
    dbeta = torch.empty(D, dtype=f32, device=x.device)
    part = _workspace(L.lr2_layernorm_bwd_partials_floats(D) * 4, x.device)
    _lib.run(L.lr2_layernorm_bwd, ptr(dy), ptr(x), ptr(gamma), ptr(stats), ptr(add), ptr(dx), ptr(dxm), ptr(dgamma),
                              ptr(dbeta), ptr(part), rows, D, float(eps), mode, g_in, g_out, g_off, float(drop_p),
                              int(seed), int(site), ptr(seed_dev), _lib.stream())
    return dx, dxm, dgamma, dbeta


# ------------------------------------------------------------- attention --
def xattn_fwd(q, k, v, H, pre_scale, post_scale):
    """q [items,Sq,E], k/v [items,Skv,E] (views with unit inner stride allowed)."""
    L = _L()
    items, Sq, E = q.shape
    Skv = k.shape[1]
    dh = E // H
    for t in (q, k, v):
        if t.dtype != bf16 or t.stride(2) != 1 or t.stride(0) != t.shape[1] * t.stride(1):
            raise _lib.Lr2Error("xattn operands must be bf16 [items,S,E] row-pitched views")
    if k.stride(1) != v.stride(1):
        raise _lib.Lr2Error("k and v must share a pitch")
    o = torch.empty((items, Sq, E), dtype=bf16, device=q.device)
    _lib.run(L.lr2_xattn_fwd, ptr(q), q.stride(1), ptr(k), ptr(v), k.stride(1), ptr(o), E, items, Sq, Skv, H, dh,
                          float(pre_scale), float(post_scale), _lib.stream())
    return o


def xattn_bwd(q, k, v, d_o, H, pre_scale, post_scale, dkv_out=None):
    L = _L()
    items, Sq, E = q.shape
    Skv = k.shape[1]
    dh = E // H
    _cuda(d_o, bf16, "d_o")
    dq = torch.empty((items, Sq, E), dtype=bf16, device=q.device)
    if dkv_out is None:
        dk = torch.empty((items, Skv, E), dtype=bf16, device=q.device)
        dv = torch.empty((items, Skv, E), dtype=bf16, device=q.device)
    else:
        dk, dv = dkv_out
    if dk.stride(1) != dv.stride(1):
        raise _lib.Lr2Error("dk and dv must share a pitch")
    _lib.run(L.lr2_xattn_bwd, ptr(q), q.stride(1), ptr(k), ptr(v), k.stride(1), ptr(d_o), E, ptr(dq), E, ptr(dk),
                          ptr(dv), dk.stride(1), items, Sq, Skv, H, dh, float(pre_scale), float(post_scale),
                          _lib.stream())
    return dq, dk, dv


# ------------------------------------------------------------------ glue --
def cast_gather(src, index=None, out=None):
    """src f32 [bs, T, ...] -> bf16 [bs, T_dst, ...] with optional int64 index [bs, T_dst]."""
    L = _L()
    _cuda(src, f32, "src"); _cuda(index, i64, "index")
    bs, T_src = src.shape[:2]
    row = src[0, 0].numel()
    T_dst = index.shape[1] if index is not None else T_src
    if out is None:
        out = torch.empty((bs, T_dst) + tuple(src.shape[2:]), dtype=bf16, device=src.device)
    _lib.run(L.lr2_cast_gather_bf16, ptr(src), ptr(index), ptr(out), bs, T_src, T_dst, row, _lib.stream())
    return out


def rows_copy(src, src_gstride, src_off, dst, dst_gstride, dst_off, groups, rows_per_group, D, accumulate=False):
    L = _L()
    _cuda(src, bf16, "src"); _cuda(dst, bf16, "dst")
    _lib.run(L.lr2_rows_copy_bf16, ptr(src), src_gstride, src_off, ptr(dst), dst_gstride, dst_off, groups,
                               rows_per_group, D, int(accumulate), _lib.stream())
    return dst


def colsum(x, out=None, accumulate=False):
    L = _L()
    if x.dtype != bf16 or x.dim() != 2 or x.stride(1) != 1:
        raise _lib.Lr2Error("colsum expects a 2-D bf16 tensor")
    rows, cols = x.shape
    if out is None:
        out = torch.empty(cols, dtype=f32, device=x.device)
    part = _workspace(L.lr2_colsum_partials_floats(cols) * 4, x.device)
    _lib.run(L.lr2_colsum_bf16, ptr(x), x.stride(0), rows, cols, ptr(out), ptr(part), int(accumulate), _lib.stream())
    return out


def rowdot_fwd(x, w, b, rows, row_stride=1, row_off=0):
    L = _L()
    _cuda(x, bf16, "x"); _cuda(w, f32, "w"); _cuda(b, f32, "b")
    D = x.shape[-1]
    out = torch.empty(rows, dtype=f32, device=x.device)
    _lib.run(L.lr2_rowdot_fwd, ptr(x), row_stride, row_off, ptr(w), ptr(b), ptr(out), rows, D, _lib.stream())
    return out


def rowdot_bwd(x, w, dout, rows, row_stride=1, row_off=0):
    L = _L()
    _cuda(x, bf16, "x"); _cuda(w, f32, "w"); _cuda(dout, f32, "dout")
    D = x.shape[-1]
    dx = torch.empty((rows * row_stride, D), dtype=bf16, device=x.device)
    dw = torch.empty(D, dtype=f32, device=x.device)
    db = torch.empty(1, dtype=f32, device=x.device)
    _lib.run(L.lr2_rowdot_bwd, ptr(x), row_stride, row_off, ptr(w), ptr(dout), ptr(dx), ptr(dw), ptr(db), rows, D,
                           _lib.stream())
    return dx, dw, db


def add_pos_fwd(x, pos, bs, T):
    L = _L()
    _cuda(x, bf16, "x"); _cuda(pos, f32, "pos")
    _lib.run(L.lr2_add_pos_fwd, ptr(x), ptr(pos), bs, T, x.shape[-1], _lib.stream())
    return x


def add_pos_bwd(dx, bs, T):
    L = _L()
    _cuda(dx, bf16, "dx")
    D = dx.shape[-1]
    dpos = torch.empty((T, D), dtype=f32, device=dx.device)
    _lib.run(L.lr2_add_pos_bwd, ptr(dx), ptr(dpos), bs, T, D, _lib.stream())
    return dpos


def to_bf16(src, out=None):
    L = _L()
    _cuda(src, f32, "src")
    if out is None:
        out = torch.empty(src.shape, dtype=bf16, device=src.device)
    _lib.run(L.lr2_cast_f32_to_bf16, ptr(src), ptr(out), src.numel(), _lib.stream())
    return out


def to_f32(src, out=None):
    L = _L()
    _cuda(src, bf16, "src")
    if out is None:
        out = torch.empty(src.shape, dtype=f32, device=src.device)
    _lib.run(L.lr2_cast_bf16_to_f32, ptr(src), ptr(out), src.numel(), _lib.stream())
    return out


# -------------------------------------------------------------- PPO rows --
def ppo_policy_loss(s, s_old, reward, v_old, pi, w_kl, w_ent, margin=0.01, adv_eps=-0.1, want_grad=True):
    L = _L()
    for t, nm in ((s, "s"), (s_old, "s_old"), (reward, "reward"), (v_old, "v_old")):
        _cuda(t, f32, nm)
    _cuda(pi, i64, "pi")
    if s.dim() != 2:
        raise _lib.Lr2Error("s must be [B, n]")
    B, n = s.shape
    if s_old.shape != s.shape:
        raise _lib.Lr2Error(f"s_old must have the shape of s {tuple(s.shape)}, got {tuple(s_old.shape)}")
    if pi.dim() != 2 or pi.shape[0] != B:
        raise _lib.Lr2Error(f"pi must be [B, k] with B = {B}, got {tuple(pi.shape)}")
    k = pi.shape[1]
    if reward.numel() != B or v_old.numel() != B:
        raise _lib.Lr2Error(f"reward and v_old must hold B = {B} values")
    dev = s.device
    scal = torch.empty(4, dtype=f32, device=dev)
    kl = torch.empty(B, dtype=f32, device=dev)
    ent = torch.empty(B, dtype=f32, device=dev)
    radj = torch.empty(B, dtype=f32, device=dev)
    adv = torch.empty(B, dtype=f32, device=dev)
    ds = torch.empty((B, n), dtype=f32, device=dev) if want_grad else None
    _lib.run(L.lr2_ppo_policy_loss, ptr(s), ptr(s_old), ptr(reward), ptr(v_old), ptr(pi), B, n, k, float(w_kl),
                                float(w_ent), float(margin), float(adv_eps), ptr(scal), ptr(kl), ptr(ent),
                                ptr(radj), ptr(adv), ptr(ds), _lib.stream())
    return dict(loss=scal[0], rank_loss=scal[1], hinge_cnt=scal[2], sum_abs_adv=scal[3], kl=kl, entropy=ent,
                reward_adj=radj, adv=adv, ds=ds)


def clipped_value_loss(v, ret, v_old, clip, want_grad=True):
    L = _L()
    for t in (v, ret, v_old):
        _cuda(t, f32)
    B = v.numel()
    loss = torch.empty(1, dtype=f32, device=v.device)
    dv = torch.empty(B, dtype=f32, device=v.device) if want_grad else None
    _lib.run(L.lr2_clipped_value_loss, ptr(v), ptr(ret), ptr(v_old), B, float(clip), ptr(loss), ptr(dv), _lib.stream())
    return loss[0], dv


def pair_hinge_loss(chosen, reject, margin=1.0, want_grad=True):
    L = _L()
    _cuda(chosen, f32); _cuda(reject, f32)
    B = chosen.numel()
    out = torch.empty(2, dtype=f32, device=chosen.device)
    dc = torch.empty(B, dtype=f32, device=chosen.device) if want_grad else None
    dr = torch.empty(B, dtype=f32, device=chosen.device) if want_grad else None
    _lib.run(L.lr2_pair_hinge_loss, ptr(chosen), ptr(reject), B, float(margin), ptr(out), ptr(dc), ptr(dr),
                                _lib.stream())
    return out[0], out[1], dc, dr


def smooth_l1_loss(logits, tgt, beta=0.3, want_grad=True):
    L = _L()
    _cuda(logits, f32); _cuda(tgt, i64)
    n = logits.numel()
    loss = torch.empty(1, dtype=f32, device=logits.device)
    dl = torch.empty(n, dtype=f32, device=logits.device) if want_grad else None
    _lib.run(L.lr2_smooth_l1_loss, ptr(logits), ptr(tgt), n, float(beta), ptr(loss), ptr(dl), _lib.stream())
    return loss[0], dl


def ppo_rollout(scores, state=None, n_prefix=2, want_order=False):
    L = _L()
    _cuda(scores, f32); _cuda(state, i64)
    B, n = scores.shape
    ns = torch.empty((B, n_prefix + n), dtype=i64, device=scores.device)
    order = torch.empty((B, n), dtype=i64, device=scores.device) if want_order else None
    _lib.run(L.lr2_ppo_rollout, ptr(scores), ptr(state), B, n, n_prefix, ptr(ns), ptr(order), _lib.stream())
    return (ns, order) if want_order else ns


def rank_sample(scores, u=None, greedy=False):
    L = _L()
    _cuda(scores, f32); _cuda(u, f32)
    B, n = scores.shape
    perm = torch.empty((B, n), dtype=i64, device=scores.device)
    lp = torch.empty(B, dtype=f32, device=scores.device)
    _lib.run(L.lr2_rank_sample, ptr(scores), ptr(u), B, n, int(greedy), ptr(perm), ptr(lp), _lib.stream())
    return perm, lp


def rank_logprob(scores, perm, dlogprob=None, want_grad=False):
    """Plackett-Luce log-probability of the rankings perm i64 [B,n] under scores f32 [B,n] -> lp f32 [B]
    (+ dscores f32 [B,n] = dlogprob * d lp / d scores when want_grad)."""
    L = _L()
    _cuda(scores, f32, "scores"); _cuda(perm, i64, "perm"); _cuda(dlogprob, f32, "dlogprob")
    B, n = scores.shape
    if perm.shape != (B, n):
        raise _lib.Lr2Error("perm must have the shape of scores")
    lp = torch.empty(B, dtype=f32, device=scores.device)
    ds = torch.empty((B, n), dtype=f32, device=scores.device) if want_grad else None
    inv = torch.empty((B, n), dtype=torch.int32, device=scores.device)
    _lib.run(L.lr2_rank_logprob, ptr(scores), ptr(perm), B, n, ptr(dlogprob), ptr(lp), ptr(ds), ptr(inv), _lib.stream())
    return (lp, ds) if want_grad else lp


def ppo_clip_surrogate(logp, logp_old, adv, eps_clip=0.2, normalize=False, norm_eps=1e-5, want_grad=True):
    """-> dict(loss, clip_frac, dlogp [B] | None, adv [B] as used)."""
    L = _L()
    for t, nm in ((logp, "logp"), (logp_old, "logp_old"), (adv, "adv")):
        _cuda(t, f32, nm)
    B = logp.numel()
    if logp_old.numel() != B or adv.numel() != B:
        raise _lib.Lr2Error("logp, logp_old and adv must have the same number of rows")
    out = torch.empty(2, dtype=f32, device=logp.device)
    dl = torch.empty(B, dtype=f32, device=logp.device) if want_grad else None
    used = torch.empty(B, dtype=f32, device=logp.device)
    _lib.run(L.lr2_ppo_clip_surrogate, ptr(logp), ptr(logp_old), ptr(adv), B, float(eps_clip), int(bool(normalize)),
                                   float(norm_eps), ptr(out), ptr(dl), ptr(used), _lib.stream())
    return dict(loss=out[0], clip_frac=out[1], dlogp=dl, adv=used)


def gae_scan(rewards, values, gamma, lam, notdone=None):
    L = _L()
    _cuda(rewards, f32); _cuda(values, f32); _cuda(notdone, f32)
    B, T = rewards.shape
    if values.shape != (B, T + 1):
        raise _lib.Lr2Error("values must be [B, T+1]")
    adv = torch.empty((B, T), dtype=f32, device=rewards.device)
    ret = torch.empty((B, T), dtype=f32, device=rewards.device)
    _lib.run(L.lr2_gae_scan, ptr(rewards), ptr(values), ptr(notdone), B, T, float(gamma), float(lam), ptr(adv), ptr(ret),
                         _lib.stream())
    return adv, ret


# ------------------------------------------------------------------ NDCG --
_log2_tables = {}


def log2_table(n, device):
    """fp32 log2(i+2), produced by the same torch call the reference makes (ndcg.py:31)."""
    key = (str(device), n)
    if key not in _log2_tables:
        tab = torch.log2(torch.arange(2, n + 2, dtype=torch.int64))  # int64 -> float32, as in the reference
        _log2_tables[key] = tab.to(device)
    return _log2_tables[key]


_ks_cache = {}


def _ks_tensor(ks, dev):
    key = (str(dev), tuple(ks))
    if key not in _ks_cache:
        _ks_cache[key] = torch.tensor(list(ks), dtype=i64, device=dev)
    return _ks_cache[key]


def ndcg_at_k(scores, labels, ks, lens=None, want_order=False):
    """scores f32 [B,N], labels i64 [B,N], ks list[int] -> ndcg f32 [B, len(ks)] (+ order i64 [B,N])."""
    L = _L()
    _cuda(scores, f32); _cuda(labels, i64)
    if lens is not None:
        _cuda(lens, torch.int32)
    B, N = scores.shape
    dev = scores.device
    ks_t = _ks_tensor(ks, dev)
    out = torch.empty((B, len(ks)), dtype=f32, device=dev)
    order = torch.full((B, N), -1, dtype=i64, device=dev) if want_order else None
    _lib.run(L.lr2_ndcg_at_k, ptr(scores), ptr(labels), ptr(lens), B, N, N, ptr(ks_t), len(ks), ptr(log2_table(N, dev)),
                          ptr(out), ptr(order), _lib.stream())
    return (out, order) if want_order else out


def ndcg_presorted(pred_rel, true_rel, ks, lens=None):
    """pred_rel, true_rel i64 [B,N] already in rank order -> ndcg f32 [B, len(ks)] (ndcg.py:54-65)."""
    L = _L()
    _cuda(pred_rel, i64); _cuda(true_rel, i64)
    B, N = pred_rel.shape
    dev = pred_rel.device
    ks_t = torch.tensor(list(ks), dtype=i64, device=dev)
    out = torch.empty((B, len(ks)), dtype=f32, device=dev)
    scratch = torch.empty(2 * B * len(ks), dtype=f32, device=dev)
    _lib.run(L.lr2_ndcg_presorted, ptr(pred_rel), ptr(true_rel), ptr(lens), B, N, ptr(ks_t), len(ks),
                               ptr(log2_table(N, dev)), ptr(out), ptr(scratch), _lib.stream())
    return out


def gather_rows(src, index):
    """src bf16 [bs, T_src, R], index i64 [bs, T_dst] -> bf16 [bs, T_dst, R]."""
    L = _L()
    _cuda(src, bf16, "src"); _cuda(index, i64, "index")
    bs, T_src, R = src.shape
    T_dst = index.shape[1]
    out = torch.empty((bs, T_dst, R), dtype=bf16, device=src.device)
    _lib.run(L.lr2_gather_rows_bf16, ptr(src), ptr(index), ptr(out), bs, T_src, T_dst, R, _lib.stream())
    return out


def bump_counter(counter, inc=1):
    """counter: int64 CUDA tensor of one element (device-resident dropout seed offset)."""
    L = _L()
    _cuda(counter, i64, "counter")
    _lib.run(L.lr2_bump_counter, ptr(counter), int(inc), _lib.stream())
    return counter


# ------------------------------------------------- TencentPretrain tower kernels --
def mha_fwd(qkv, B, S, H, key_bias=None, scale=None, drop_p=0.0, seed=0, seed_dev=None, want_lse=True):
    """qkv: bf16 [B*S, 3*E] merged projection output (q | k | v).  Returns (o [B*S, E], lse [B,H,S])."""
    L = _L()
    _cuda(qkv, bf16, "qkv"); _cuda(key_bias, f32, "key_bias")
    E = qkv.shape[1] // 3
    dh = E // H
    o = torch.empty((B * S, E), dtype=bf16, device=qkv.device)
    lse = torch.empty((B, H, S), dtype=f32, device=qkv.device) if want_lse else None
    sc = 1.0 / math.sqrt(dh) if scale is None else scale
    _lib.run(L.lr2_mha_fwd, ptr(qkv), qkv.data_ptr() + 2 * E, qkv.data_ptr() + 4 * E, qkv.stride(0), ptr(key_bias),
             ptr(o), E, ptr(lse), B, S, H, dh, float(sc), float(drop_p), int(seed), ptr(seed_dev), _lib.stream())
    return o, lse


def mha_bwd(qkv, o, d_o, lse, B, S, H, key_bias=None, scale=None, drop_p=0.0, seed=0, seed_dev=None):
    """Returns dqkv bf16 [B*S, 3*E] (dq | dk | dv)."""
    L = _L()
    _cuda(qkv, bf16); _cuda(o, bf16); _cuda(d_o, bf16); _cuda(lse, f32)
    E = qkv.shape[1] // 3
    dh = E // H
    d = torch.empty_like(qkv)
    sc = 1.0 / math.sqrt(dh) if scale is None else scale
    _lib.run(L.lr2_mha_bwd, ptr(qkv), qkv.data_ptr() + 2 * E, qkv.data_ptr() + 4 * E, qkv.stride(0), ptr(key_bias),
             ptr(o), ptr(d_o), E, ptr(lse), ptr(d), d.data_ptr() + 2 * E, d.data_ptr() + 4 * E, d.stride(0), B, S, H, dh,
             float(sc), float(drop_p), int(seed), ptr(seed_dev), _lib.stream())
    return d


def embed_sum(src, seg, word, pos, seg_table=None):
    L = _L()
    _cuda(src, i64); _cuda(word, f32); _cuda(pos, f32); _cuda(seg_table, f32)
    B, S = src.shape
    D = word.shape[1]
    out = torch.empty((B * S, D), dtype=bf16, device=src.device)
    segp = seg.to(i64).contiguous() if seg_table is not None else None
    _lib.run(L.lr2_embed_sum, ptr(src.contiguous()), ptr(segp), ptr(word), ptr(pos), ptr(seg_table), ptr(out), B * S, S,
             D, _lib.stream())
    return out


def embed_scatter_add(idx, d, table):
    L = _L()
    _cuda(idx, i64); _cuda(d, bf16); _cuda(table, f32)
    _lib.run(L.lr2_embed_scatter_add, ptr(idx.contiguous()), ptr(d), ptr(table), d.shape[0], d.shape[1], _lib.stream())
    return table


def patchify(img, ps):
    L = _L()
    _cuda(img, f32, "img")
    B, C, Hh, Ww = img.shape
    out = torch.empty((B * (Hh // ps) * (Ww // ps), C * ps * ps), dtype=bf16, device=img.device)
    _lib.run(L.lr2_patchify, ptr(img), ptr(out), B, C, Hh, Ww, ps, _lib.stream())
    return out


def bias_gelu_rows(x, bias, want_pre=True):
    """x f32 [rows, D] -> (gelu(bf16(x + bias)) bf16, bf16(x + bias) | None)."""
    L = _L()
    _cuda(x, f32, "x"); _cuda(bias, f32, "bias")
    rows, D = x.shape
    out = torch.empty((rows, D), dtype=bf16, device=x.device)
    pre = torch.empty((rows, D), dtype=bf16, device=x.device) if want_pre else None
    _lib.run(L.lr2_bias_gelu_rows, ptr(x), ptr(bias), ptr(out), ptr(pre), rows, D, _lib.stream())
    return out, pre


def dropout(x, p, seed, site, seed_dev=None):
    L = _L()
    _cuda(x, bf16)
    out = torch.empty_like(x)
    _lib.run(L.lr2_dropout_bf16, ptr(x), ptr(out), x.numel(), float(p), int(seed), int(site), ptr(seed_dev),
             _lib.stream())
    return out
