"""CLIP-style VideoTransformer with the reference's constructor and state_dict keys
(finetune/video_transformer.py:8-93): class token + positional embedding -> ln_pre -> `layers` pre-LN residual
attention blocks (nn.MultiheadAttention parameters: in_proj_weight / in_proj_bias / out_proj; QuickGELU MLP) ->
ln_post -> @ proj.  It is imported but never instantiated by the reference's scripts, so this is
API parity: the blocks reuse the tower layer engine (merged-QKV GEMM, lr2_mha attention, QuickGELU epilogues).
Requires emb_size / heads == 64 and frame_size + 1 <= 256 (the attention kernel's limits)."""
import math

import torch
import torch.nn as nn

from . import engine as eng
from . import ops, tower
from .ops import EPI_BIAS_QGELU, EPI_DQGELU

bf16 = torch.bfloat16


class LayerNorm(nn.LayerNorm):
    """fp32 LayerNorm of the reference (statistics are always computed in fp32 by the kernel)."""


class QuickGELU(nn.Module):
    pass


class ResidualAttentionBlock(nn.Module):
    def __init__(self, d_model, n_head, attn_mask=None):
        super().__init__()
        if attn_mask is not None:
            raise ValueError("attn_mask is None wherever the reference builds this block")
        self.attn = nn.MultiheadAttention(d_model, n_head)
        self.ln_1 = LayerNorm(d_model)
        self.mlp = nn.Sequential()
        self.mlp.add_module("c_fc", nn.Linear(d_model, d_model * 4))
        self.mlp.add_module("gelu", QuickGELU())
        self.mlp.add_module("c_proj", nn.Linear(d_model * 4, d_model))
        self.ln_2 = LayerNorm(d_model)
        self.attn_mask = attn_mask
        self.n_head = n_head


class Transformer(nn.Module):
    def __init__(self, width, layers, heads, attn_mask=None):
        super().__init__()
        self.width, self.layers = width, layers
        self.resblocks = nn.Sequential(*[ResidualAttentionBlock(width, heads, attn_mask) for _ in range(layers)])


class _BlockW:
    """Adapter: exposes a ResidualAttentionBlock to tower.layer_forward / layer_backward (pre-LN, no dropout)."""

    def __init__(self, bank, blk):
        a = blk.attn
        E = a.embed_dim
        self.wqkv, self.bqkv = bank.get(a.in_proj_weight), a.in_proj_bias.detach()
        self.qkv_targets = [(a.in_proj_weight, a.in_proj_bias, 0, 3 * E)]
        self.lin_o, self.lin_1, self.lin_2 = a.out_proj, blk.mlp.c_fc, blk.mlp.c_proj
        self.wo, self.bo = bank.get(a.out_proj.weight), a.out_proj.bias.detach()
        self.w1, self.b1 = bank.get(blk.mlp.c_fc.weight), blk.mlp.c_fc.bias.detach()
        self.w2, self.b2 = bank.get(blk.mlp.c_proj.weight), blk.mlp.c_proj.bias.detach()
        self.ln1, self.ln2 = blk.ln_1, blk.ln_2
        self.heads = blk.n_head
        self.scale = 1.0 / math.sqrt(E // blk.n_head)
        self.p_att = self.p1 = self.p2 = 0.0
        self.pre = True
        self.epi_act, self.epi_dact = EPI_BIAS_QGELU, EPI_DQGELU


class _VTFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, x, save, *params):
        B, Fr, E = x.shape
        S = Fr + 1
        xb = x.detach()
        xb = xb.contiguous() if xb.dtype == bf16 else ops.to_bf16(xb.float().contiguous())
        h0 = torch.empty((B * S, E), dtype=bf16, device=x.device)
        ops.rows_copy(xb.view(B * Fr, E), Fr, 0, h0, S, 1, B, Fr, E)
        h0.view(B, S, E)[:, 0, :] = mod.class_embedding.detach().to(bf16)
        ops.add_pos_fwd(h0, mod.positional_embedding.detach().contiguous(), B, S)
        h, st_pre = tower._ln(h0, mod.ln_pre, save)
        ctxs, Ws = [], []
        for i, blk in enumerate(mod.transformer.resblocks):
            W = _BlockW(mod._bank, blk)
            h, c = tower.layer_forward(W, h, None, B, S, False, 0, None, 10 + 2 * i, save)
            ctxs.append(c); Ws.append(W)
        y, st_post = tower._ln(h, mod.ln_post, save)
        out = ops.gemm(y, mod._bank.get(mod.proj), b_mn=True)               # x @ proj, proj stored [E, out]
        ctx.saved = (h0, st_pre, ctxs, Ws, h, st_post, y, B, S, E, Fr) if save else None
        ctx.mod, ctx.np, ctx.dt = mod, len(params), x.dtype
        out = out.view(B, S, -1)
        return out if x.dtype == bf16 else ops.to_f32(out)

    @staticmethod
    def backward(ctx, dout):
        mod = ctx.mod
        h0, st_pre, ctxs, Ws, h_last, st_post, y, B, S, E, Fr = ctx.saved
        d = dout.reshape(B * S, -1)
        d = d.contiguous() if d.dtype == bf16 else ops.to_bf16(d.float().contiguous())
        sink = eng._GradSink()
        sink.put_vec(mod.proj, ops.gemm(y, d, a_mn=True, b_mn=True, out_dtype=torch.float32))   # y^T d
        dy = ops.gemm(d, mod._bank.get(mod.proj))                                               # d @ proj^T
        dh, _ = tower._ln_bwd(sink, mod.ln_post, dy, h_last, st_post)
        dm = None
        for i in range(len(ctxs) - 1, -1, -1):
            dh, dm = tower.layer_backward(Ws[i], ctxs[i], dh, dm, sink)
        d0, _ = tower._ln_bwd(sink, mod.ln_pre, dh, h0, st_pre)
        dpos = ops.add_pos_bwd(d0, B, S)
        sink.put_vec(mod.positional_embedding, dpos)
        sink.put_vec(mod.class_embedding, dpos[0])
        dx = torch.empty((B * Fr, E), dtype=bf16, device=d.device)
        ops.rows_copy(d0, S, 1, dx, Fr, 0, B, Fr, E)
        dx = dx.view(B, Fr, E)
        return (None, dx if ctx.dt == bf16 else ops.to_f32(dx), None) + (None,) * ctx.np


class VideoTransformer(nn.Module):
    def __init__(self, frame_size, emb_size, layers, heads, output_dim):
        super().__init__()
        self.emb_size, self.output_dim, self.frame_size = emb_size, output_dim, frame_size
        if emb_size // heads != 64 or frame_size + 1 > 256:
            raise ValueError("lr2ppo_b200.VideoTransformer needs emb_size/heads == 64 and frame_size + 1 <= 256")
        scale = emb_size ** -0.5
        self.class_embedding = nn.Parameter(scale * torch.randn(emb_size))
        self.positional_embedding = nn.Parameter(scale * torch.randn(frame_size + 1, emb_size))
        self.ln_pre = LayerNorm(emb_size)
        self.transformer = Transformer(emb_size, layers, heads)
        self.ln_post = LayerNorm(emb_size)
        self.proj = nn.Parameter(scale * torch.randn(emb_size, output_dim))
        self._bank = tower._Bank()

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("lr2ppo_b200.VideoTransformer runs on CUDA (sm_100a) only; there is no CPU fallback")
        if x.shape[1] != self.frame_size:
            raise ValueError("expected [N, frame_size, emb_size] input")
        params = list(self.parameters())
        save = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in params))
        return _VTFn.apply(self, x, save, *params)
