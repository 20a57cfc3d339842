"""TencentPretrain encoder towers (ViT-B/16 and RoBERTa-base as configured by models/vit/base-16-224_config.json
and models/xlm-roberta/base_config.json) on the sm_100a kernels, with the reference's module names, constructor
arguments and state_dict keys:

  LayerNorm                tencentpretrain/layers/layer_norm.py:5-21     (gamma/beta, unbiased std + eps)
  MultiHeadedAttention     tencentpretrain/layers/multi_headed_attn.py:6-76
  PositionwiseFeedForward  tencentpretrain/layers/position_ffn.py:4-15
  TransformerLayer         tencentpretrain/layers/transformer.py:8-73    (post-LN and pre-LN)
  TransformerEncoder       tencentpretrain/encoders/transformer_encoder.py:7-138 ("fully_visible" mask)
  Embedding / PatchEmbedding / WordEmbedding / PosEmbedding / SegEmbedding   tencentpretrain/embeddings/*.py
  Model, build_model       tencentpretrain/models/model.py, model_builder.py:8-49 (embedding + encoder; targets are
                           host-side heads outside the kernel scope and are not built)

The sub-modules hold parameters; the computation runs in `EncoderEngine` (explicit forward / backward over bf16
activations: merged-QKV tcgen05 GEMMs with fused bias / GELU / dropout / residual epilogues, the flash-style
`lr2_mha` attention kernel, the TencentPretrain LayerNorm kernel).  Out of scope, as in SURVEY §2 row 10:
relative position bias, residual attention, gated FFN, T5 LayerNorm, parameter sharing, causal masks.
"""
import math

import torch
import torch.nn as nn

from . import engine as eng
from . import ops
from .ops import EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_DROP_RES, EPI_DGELU, EPI_ADD

bf16 = torch.bfloat16
LN_MODE = 1  # TencentPretrain LayerNorm


class LayerNorm(nn.Module):
    def __init__(self, hidden_size, eps=1e-6):
        super().__init__()
        self.eps = eps
        self.gamma = nn.Parameter(torch.ones(hidden_size))
        self.beta = nn.Parameter(torch.zeros(hidden_size))


class MultiHeadedAttention(nn.Module):
    def __init__(self, hidden_size, heads_num, attention_head_size, dropout, has_bias=True, with_scale=True):
        super().__init__()
        self.heads_num = heads_num
        self.per_head_size = attention_head_size
        self.with_scale = with_scale
        self.inner_hidden_size = heads_num * attention_head_size
        self.linear_layers = nn.ModuleList([nn.Linear(hidden_size, self.inner_hidden_size, bias=has_bias)
                                            for _ in range(3)])
        self.dropout = nn.Dropout(dropout)
        self.final_linear = nn.Linear(self.inner_hidden_size, hidden_size, bias=has_bias)
        if not has_bias:
            raise ValueError("remove_transformer_bias is not used by the LR2PPO tower configs")


class PositionwiseFeedForward(nn.Module):
    def __init__(self, hidden_size, feedforward_size, hidden_act, has_bias=True):
        super().__init__()
        if hidden_act != "gelu":
            raise ValueError("the LR2PPO tower configs use hidden_act = gelu")
        self.linear_1 = nn.Linear(hidden_size, feedforward_size, bias=has_bias)
        self.linear_2 = nn.Linear(feedforward_size, hidden_size, bias=has_bias)


def _arg(args, name, default):
    return getattr(args, name, default)


class TransformerLayer(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.layernorm_positioning = _arg(args, "layernorm_positioning", "post")
        head = _arg(args, "attention_head_size", args.hidden_size // args.heads_num)
        if _arg(args, "feed_forward", "dense") != "dense" or _arg(args, "layernorm", "normal") != "normal":
            raise ValueError("gated FFN / T5 LayerNorm are outside the LR2PPO hot path")
        self.self_attn = MultiHeadedAttention(args.hidden_size, args.heads_num, head, args.dropout,
                                              has_bias=not _arg(args, "remove_transformer_bias", False),
                                              with_scale=not _arg(args, "remove_attention_scale", False))
        self.dropout_1 = nn.Dropout(args.dropout)
        self.feed_forward = PositionwiseFeedForward(args.hidden_size, args.feedforward_size, args.hidden_act)
        self.dropout_2 = nn.Dropout(args.dropout)
        self.layer_norm_1 = LayerNorm(args.hidden_size)
        self.layer_norm_2 = LayerNorm(args.hidden_size)


# ----------------------------------------------------------------------------------------- engine ---------
class _LayerW:
    """bf16 shadows of one layer: merged [3E, E] QKV weight (views are the per-Linear shadows) + fp32 biases."""

    def __init__(self, bank, layer):
        att, ff = layer.self_attn, layer.feed_forward
        E = att.linear_layers[0].weight.shape[1]
        inner = att.inner_hidden_size
        key = id(layer)
        merged = bank.extra.get(key)
        if merged is None:
            merged = torch.empty((3 * inner, E), dtype=bf16, device=att.final_linear.weight.device)
            bank.extra[key] = merged
        for i, lin in enumerate(att.linear_layers):
            bank.get_into(lin.weight, merged[i * inner:(i + 1) * inner])
        self.wqkv = merged
        self.bqkv = torch.cat([lin.bias.detach() for lin in att.linear_layers])
        self.qkv_targets = [(lin.weight, lin.bias, i * inner, (i + 1) * inner) for i, lin in enumerate(att.linear_layers)]
        self.lin_o, self.lin_1, self.lin_2 = att.final_linear, ff.linear_1, ff.linear_2
        self.epi_act, self.epi_dact = EPI_BIAS_GELU, EPI_DGELU
        self.att, self.ff, self.layer = att, ff, layer
        self.wo, self.bo = bank.get(att.final_linear.weight), att.final_linear.bias.detach()
        self.w1, self.b1 = bank.get(ff.linear_1.weight), ff.linear_1.bias.detach()
        self.w2, self.b2 = bank.get(ff.linear_2.weight), ff.linear_2.bias.detach()
        self.ln1, self.ln2 = layer.layer_norm_1, layer.layer_norm_2
        self.heads = att.heads_num
        self.scale = 1.0 / math.sqrt(float(att.per_head_size)) if att.with_scale else 1.0
        self.p_att = att.dropout.p
        self.p1, self.p2 = layer.dropout_1.p, layer.dropout_2.p
        self.pre = layer.layernorm_positioning == "pre"


def _lnp(ln):
    """(gamma, beta, eps, kernel mode) of a TencentPretrain LayerNorm (gamma/beta) or a torch nn.LayerNorm."""
    if hasattr(ln, "gamma"):
        return ln.gamma, ln.beta, ln.eps, LN_MODE
    return ln.weight, ln.bias, ln.eps, 0


def _ln(x, ln, save, out=None):
    g, b, eps, mode = _lnp(ln)
    return ops.layernorm_fwd(x, g.detach(), b.detach(), eps, mode, out=out, want_stats=save)


def _ln_bwd(sink, ln, dy, x, st, **kw):
    g, b, eps, mode = _lnp(ln)
    dx, dxm, dg, db = ops.layernorm_bwd(dy, x, g.detach(), st, eps, mode, **kw)
    sink.put_vec(g, dg); sink.put_vec(b, db)
    return dx, dxm


def layer_forward(W, h, kbias, B, S, train, seed, seed_dev, site, save):
    """One TransformerLayer (ref: layers/transformer.py:50-73).  h: bf16 [B*S, E]."""
    pa = W.p_att if train else 0.0
    p1 = W.p1 if train else 0.0
    p2 = W.p2 if train else 0.0
    c = {}
    if W.pre:
        l1, st1 = _ln(h, W.ln1, save)
        a_in = l1
    else:
        a_in, st1 = h, None
    qkv = ops.gemm(a_in, W.wqkv, epilogue=EPI_BIAS, bias=W.bqkv)
    a, lse = ops.mha_fwd(qkv, B, S, W.heads, kbias, W.scale, pa, seed + 7919 * site, seed_dev, want_lse=save)
    t = ops.gemm(a, W.wo, epilogue=EPI_BIAS_DROP_RES, bias=W.bo, aux=h, drop_p=p1, seed=seed, site=site,
                 seed_dev=seed_dev)                                   # dropout_1(attn) + hidden
    if W.pre:
        l2, st2 = _ln(t, W.ln2, save)
        f_in, res2 = l2, t
    else:
        i1, st2 = _ln(t, W.ln1, save)                                 # post-LN: layer_norm_1(inter + hidden)
        f_in, res2 = i1, i1
    pre = torch.empty((h.shape[0], W.w1.shape[0]), dtype=bf16, device=h.device) if save else None
    f = ops.gemm(f_in, W.w1, epilogue=W.epi_act, bias=W.b1, c2=pre)
    t2 = ops.gemm(f, W.w2, epilogue=EPI_BIAS_DROP_RES, bias=W.b2, aux=res2, drop_p=p2, seed=seed, site=site + 1,
                  seed_dev=seed_dev)                                  # dropout_2(ffn) + residual
    if W.pre:
        out, st3 = t2, None
    else:
        out, st3 = _ln(t2, W.ln2, save)                               # post-LN: layer_norm_2(output + inter)
    if save:
        c = dict(h=h, a_in=a_in, st1=st1, qkv=qkv, a=a, lse=lse, t=t, f_in=f_in, st2=st2, pre=pre, f=f, t2=t2,
                 st3=st3, p=(pa, p1, p2), seed=seed, seed_dev=seed_dev, site=site, B=B, S=S, kbias=kbias)
    return out, c


def _put_lin(sink, lin, dy, x):
    eng._wgrad(sink, lin, dy, x)


def layer_backward(W, c, dout, dout_m, sink, prev_site2=None, prev_p2=0.0):
    """Backward of layer_forward.
    pre-LN : dout = grad wrt the layer output h3 = drop2(ffn) + h2; dout_m = dout with drop2's mask applied.
             Returns (dh, dh_m) where dh_m carries the PREVIOUS layer's dropout_2 mask (prev_site2).
    post-LN: dout = grad wrt out = LN2(t2); dout_m unused. Returns (dh, None)."""
    pa, p1, p2 = c["p"]
    seed, sdev, site, B, S = c["seed"], c["seed_dev"], c["site"], c["B"], c["S"]
    if W.pre:
        dy2 = dout_m if dout_m is not None else dout                     # grad into dropout_2's input
        res_grad = dout
    else:
        dt2, dy2 = _ln_bwd(sink, W.ln2, dout, c["t2"], c["st3"], drop_p=p2, seed=seed, site=site + 1,
                           want_masked=True, seed_dev=sdev)
        res_grad = dt2
    _put_lin(sink, W.lin_2, dy2, c["f"])
    dfp = eng._dgrad(dy2, W.w2, epilogue=W.epi_dact, aux=c["pre"])
    _put_lin(sink, W.lin_1, dfp, c["f_in"])
    if W.pre:
        dl2 = eng._dgrad(dfp, W.w1)
        dt, dy1 = _ln_bwd(sink, W.ln2, dl2, c["t"], c["st2"], add=res_grad, drop_p=p1, seed=seed, site=site,
                          want_masked=True, seed_dev=sdev)
    else:
        di1 = eng._dgrad(dfp, W.w1, epilogue=EPI_ADD, aux=res_grad)     # + residual branch of i1
        dt, dy1 = _ln_bwd(sink, W.ln1, di1, c["t"], c["st2"], drop_p=p1, seed=seed, site=site, want_masked=True,
                          seed_dev=sdev)
    # attention output projection, core, merged QKV projection
    _put_lin(sink, W.lin_o, dy1, c["a"])
    da = eng._dgrad(dy1, W.wo)
    dqkv = ops.mha_bwd(c["qkv"], c["a"], da, c["lse"], B, S, W.heads, c["kbias"], W.scale, pa, seed + 7919 * site, sdev)
    gw = ops.gemm(dqkv, c["a_in"], a_mn=True, b_mn=True, out_dtype=torch.float32)          # [3*inner, E]
    gb = ops.colsum(dqkv)
    for wp, bp, lo, hi in W.qkv_targets:
        sink.put_vec(wp, gw[lo:hi])
        sink.put_vec(bp, gb[lo:hi])
    if W.pre:
        dl1 = eng._dgrad(dqkv, W.wqkv)
        return _ln_bwd(sink, W.ln1, dl1, c["h"], c["st1"], add=dt, drop_p=prev_p2, seed=seed, site=(prev_site2 or 0),
                       want_masked=prev_site2 is not None, seed_dev=sdev)
    dh = eng._dgrad(dqkv, W.wqkv, epilogue=EPI_ADD, aux=dt)
    return dh, None


class _Bank(eng.ShadowBank):
    def __init__(self):
        super().__init__()
        self.extra = {}

    def get_into(self, p, dst):
        """Like get(), but the bf16 copy lives in the caller-provided view `dst`."""
        ent = self._sh.get(id(p))
        if ent is None or ent[1] != p._version or ent[2] != p.data_ptr() or ent[0].data_ptr() != dst.data_ptr():
            tmp = ops.to_bf16(p.detach().contiguous())
            dst.copy_(tmp)
            self._sh[id(p)] = (dst, p._version, p.data_ptr())
        return dst


class TransformerEncoder(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.mask = args.mask
        self.layers_num = args.layers_num
        self.layernorm_positioning = _arg(args, "layernorm_positioning", "post")
        for flag in ("parameter_sharing", "factorized_embedding_parameterization", "relative_position_embedding",
                     "has_residual_attention"):
            if _arg(args, flag, False):
                raise ValueError(f"{flag} is outside the LR2PPO hot path (not used by the ViT / RoBERTa configs)")
        if self.mask != "fully_visible":
            raise ValueError("only mask = fully_visible is used by the LR2PPO tower configs")
        self.transformer = nn.ModuleList([TransformerLayer(args) for _ in range(self.layers_num)])
        if self.layernorm_positioning == "pre":
            self.layer_norm = LayerNorm(args.hidden_size)
        self._bank = _Bank()
        self._seed_counter = None

    # ---- explicit forward / backward -----------------------------------------------------------------------
    def _run_forward(self, emb, seg, train, save):
        """emb bf16 [B, S, E]; returns (hidden bf16 [B,S,E], ctx)."""
        B, S, E = emb.shape
        kbias = ((seg > 0).to(torch.float32) - 1.0) * 10000.0            # (1 - mask) * -10000, per key
        kbias = kbias.contiguous()
        seed = (torch.initial_seed() * 1000003 + id(self) % 9973) & 0x7FFFFFFFFFFFFFFF
        sdev = None
        if train:
            if self._seed_counter is None or self._seed_counter.device != emb.device:
                self._seed_counter = torch.zeros(1, dtype=torch.int64, device=emb.device)
            ops.bump_counter(self._seed_counter, 1)
            sdev = self._seed_counter.clone() if save else self._seed_counter
        h = emb.reshape(B * S, E)
        ctxs, Ws = [], []
        for i, layer in enumerate(self.transformer):
            W = _LayerW(self._bank, layer)
            h, c = layer_forward(W, h, kbias, B, S, train, seed, sdev, 10 + 2 * i, save)
            ctxs.append(c); Ws.append(W)
        fin = None
        if self.layernorm_positioning == "pre":
            x_last = h
            h, st = _ln(h, self.layer_norm, save)
            fin = (x_last, st)
        return h.view(B, S, E), (dict(ctxs=ctxs, Ws=Ws, fin=fin, seed=seed, sdev=sdev, B=B, S=S) if save else None)

    def _run_backward(self, ctx, dhid):
        """dhid bf16 [B*S, E] -> grad wrt emb [B*S, E]; parameter grads go to .grad."""
        sink = eng._GradSink()
        ctxs, Ws = ctx["ctxs"], ctx["Ws"]
        n = len(ctxs)
        d, dm = dhid, None
        if self.layernorm_positioning == "pre":
            x_last, st = ctx["fin"]
            last = ctxs[-1]
            d, dm = _ln_bwd(sink, self.layer_norm, dhid, x_last, st, drop_p=last["p"][2], seed=ctx["seed"],
                            site=last["site"] + 1, want_masked=True, seed_dev=ctx["sdev"])
        for i in range(n - 1, -1, -1):
            prev = ctxs[i - 1] if i > 0 else None
            d, dm = layer_backward(Ws[i], ctxs[i], d, dm, sink,
                                   prev_site2=(prev["site"] + 1) if prev is not None else None,
                                   prev_p2=prev["p"][2] if prev is not None else 0.0)
        return d

    def forward(self, emb, seg):
        if not emb.is_cuda:
            raise RuntimeError("lr2ppo_b200 TransformerEncoder runs on CUDA (sm_100a) only; there is no CPU fallback")
        params = list(self.parameters())
        save = torch.is_grad_enabled() and (emb.requires_grad or any(p.requires_grad for p in params))
        return _EncoderFn.apply(self, emb, seg, self.training, save, *params)


class _EncoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, enc, emb, seg, train, save, *params):
        e = emb.detach()
        e = e.contiguous() if e.dtype == bf16 else ops.to_bf16(e.float().contiguous())
        out, c = enc._run_forward(e, seg, train, save)
        ctx.enc, ctx.c, ctx.np, ctx.dt, ctx.shape = enc, c, len(params), emb.dtype, emb.shape
        return out if emb.dtype == bf16 else ops.to_f32(out)

    @staticmethod
    def backward(ctx, dout):
        d = dout.reshape(-1, dout.shape[-1])
        d = d.contiguous() if d.dtype == bf16 else ops.to_bf16(d.float().contiguous())
        demb = ctx.enc._run_backward(ctx.c, d).view(ctx.shape)
        ctx.c = None
        if ctx.dt != bf16:
            demb = ops.to_f32(demb)
        return (None, demb, None, None, None) + (None,) * ctx.np


# ------------------------------------------------------------------------------------- embeddings ---------
class PatchEmbedding(nn.Module):
    def __init__(self, args, _=None):
        super().__init__()
        self.cls_emb = nn.Parameter(torch.zeros(1, 1, args.emb_size))
        self.image_height, self.image_width = args.image_height, args.image_width
        self.patch_size = args.patch_size
        ch = _arg(args, "channels_num", 3)
        self.projection = nn.Conv2d(ch, args.emb_size, kernel_size=(args.patch_size,) * 2,
                                    stride=(args.patch_size,) * 2, bias=False)


class PosEmbedding(nn.Module):
    def __init__(self, args, _=None):
        super().__init__()
        self.max_seq_length = args.max_seq_length
        self.embedding = nn.Embedding(self.max_seq_length, args.emb_size)


class WordEmbedding(nn.Module):
    def __init__(self, args, vocab_size):
        super().__init__()
        self.embedding = nn.Embedding(vocab_size, args.emb_size)
        self.emb_size = args.emb_size


class SegEmbedding(nn.Module):
    def __init__(self, args, _=None):
        super().__init__()
        self.embedding = nn.Embedding(3, args.emb_size)


str2embedding = {"word": WordEmbedding, "pos": PosEmbedding, "seg": SegEmbedding, "patch": PatchEmbedding}
str2encoder = {"transformer": TransformerEncoder}


class Embedding(nn.Module):
    """Sum of the configured embeddings -> (TencentPretrain LayerNorm) -> dropout (embeddings/embedding.py:19-34).
    Supported combinations (the two LR2PPO tower configs): ["patch", "pos"] and ["word", "pos", "seg"]."""

    def __init__(self, args):
        super().__init__()
        self.embedding_name_list = []
        self.dropout = nn.Dropout(args.dropout)
        self.remove_embedding_layernorm = _arg(args, "remove_embedding_layernorm", False)
        if not self.remove_embedding_layernorm:
            self.layer_norm = LayerNorm(args.emb_size)
        self._bank = _Bank()
        self._seed_counter = None

    def update(self, embedding, embedding_name):
        setattr(self, embedding_name, embedding)
        self.embedding_name_list.append(embedding_name)

    def forward(self, src, seg):
        params = list(self.parameters())
        save = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        return _EmbeddingFn.apply(self, src, seg, self.training, save, *params)

    # explicit forward / backward
    def _run_forward(self, src, seg, train, save):
        names = self.embedding_name_list
        if names[0] == "patch":
            pe = self.patch
            B = src.shape[0]
            ps = pe.patch_size
            if src.shape[2] != pe.image_height or src.shape[3] != pe.image_width:
                raise ValueError(f"Input image size ({src.shape[2]}*{src.shape[3]}) doesn't match model "
                                 f"({pe.image_height}*{pe.image_width}).")
            patches = ops.patchify(src.float().contiguous(), ps)                     # [B*P, 3*ps*ps]
            w = self._bank.get(pe.projection.weight).view(pe.projection.weight.shape[0], -1)
            proj = ops.gemm(patches, w)                                              # Conv2d(k=s=16) as a GEMM
            P = proj.shape[0] // B
            S, E = P + 1, proj.shape[1]
            emb = torch.empty((B * S, E), dtype=bf16, device=src.device)
            ops.rows_copy(proj, P, 0, emb, S, 1, B, P, E)
            emb.view(B, S, E)[:, 0, :] = pe.cls_emb.detach().view(1, E).to(bf16)
            ops.add_pos_fwd(emb, self.pos.embedding.weight.detach()[:S].contiguous(), B, S)
            c = dict(kind="patch", patches=patches, B=B, S=S, E=E)
        else:
            B, S = src.shape
            segt = self.seg.embedding.weight.detach() if "seg" in names else None
            emb = ops.embed_sum(src, seg, self.word.embedding.weight.detach(),
                                self.pos.embedding.weight.detach(), segt)
            E = emb.shape[1]
            c = dict(kind="word", src=src, seg=seg, B=B, S=S, E=E)
        if not self.remove_embedding_layernorm:
            x = emb
            emb, st = _ln(x, self.layer_norm, save)
            c["ln"] = (x, st)
        p = self.dropout.p if train else 0.0
        if p > 0.0:
            seed = (torch.initial_seed() * 1000003 + id(self) % 9973) & 0x7FFFFFFFFFFFFFFF
            if self._seed_counter is None or self._seed_counter.device != emb.device:
                self._seed_counter = torch.zeros(1, dtype=torch.int64, device=emb.device)
            ops.bump_counter(self._seed_counter, 1)
            sdev = self._seed_counter.clone() if save else self._seed_counter
            emb = ops.dropout(emb, p, seed, 9, sdev)
            c["drop"] = (p, seed, sdev)
        return emb.view(c["B"], c["S"], c["E"]), (c if save else None)

    def _run_backward(self, c, d):
        """d: bf16 [B*S, E] gradient wrt the embedding output."""
        sink = eng._GradSink()
        B, S, E = c["B"], c["S"], c["E"]
        if "drop" in c:
            p, seed, sdev = c["drop"]
            d = ops.dropout(d, p, seed, 9, sdev)
        if "ln" in c:
            x, st = c["ln"]
            d, _ = _ln_bwd(sink, self.layer_norm, d, x, st)
        dpos = ops.add_pos_bwd(d, B, S)                                               # sum over the batch
        gpos = torch.zeros_like(self.pos.embedding.weight)
        gpos[:S] = dpos
        sink.put_vec(self.pos.embedding.weight, gpos)
        if c["kind"] == "patch":
            pe = self.patch
            sink.put_vec(pe.cls_emb, dpos[0])
            P = S - 1
            dproj = torch.empty((B * P, E), dtype=bf16, device=d.device)
            ops.rows_copy(d, S, 1, dproj, P, 0, B, P, E)
            gw = ops.gemm(dproj, c["patches"], a_mn=True, b_mn=True, out_dtype=torch.float32)
            sink.put_vec(pe.projection.weight, gw.view_as(pe.projection.weight))
        else:
            gword = torch.zeros_like(self.word.embedding.weight)
            ops.embed_scatter_add(c["src"].reshape(-1), d, gword)
            sink.put_vec(self.word.embedding.weight, gword)
            if "seg" in self.embedding_name_list:
                gseg = torch.zeros_like(self.seg.embedding.weight)
                ops.embed_scatter_add(c["seg"].reshape(-1).to(torch.int64), d, gseg)
                sink.put_vec(self.seg.embedding.weight, gseg)


class _EmbeddingFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, src, seg, train, save, *params):
        out, c = mod._run_forward(src, seg, train, save)
        ctx.mod, ctx.c, ctx.np = mod, c, len(params)
        return out

    @staticmethod
    def backward(ctx, dout):
        d = dout.reshape(-1, dout.shape[-1])
        d = d.contiguous() if d.dtype == bf16 else ops.to_bf16(d.float().contiguous())
        ctx.mod._run_backward(ctx.c, d)
        ctx.c = None
        return (None, None, None, None, None) + (None,) * ctx.np


class Model(nn.Module):
    """embedding -> encoder (tencentpretrain/models/model.py:32-41 without decoder / target)."""

    def __init__(self, args, embedding, encoder, target=None):
        super().__init__()
        self.embedding, self.encoder, self.target = embedding, encoder, target

    def forward(self, src, tgt, seg):
        emb = self.embedding(src, seg)
        return self.encoder(emb, seg).float()        # fp32 hidden states, as the reference returns


def build_model(args, vocab_size=None):
    """ref: tencentpretrain/model_builder.py:8-49.  `len(args.tokenizer.vocab)` is used when vocab_size is None."""
    if vocab_size is None:
        vocab_size = len(args.tokenizer.vocab)
    embedding = Embedding(args)
    for name in args.embedding:
        if name not in str2embedding:
            raise ValueError(f"embedding {name!r} is outside the LR2PPO hot path")
        embedding.update(str2embedding[name](args, vocab_size), name)
    encoder = str2encoder[args.encoder](args)
    return Model(args, embedding, encoder)
