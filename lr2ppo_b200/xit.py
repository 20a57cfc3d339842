"""XiT cross-modal fusion block with the reference's module tree (finetune/xit.py:9-148), so that
`state_dict()` keys (`0.0.0.fn.0.ln_x.weight`, `0.0.0.fn.1.queries.weight`, `0.0.1.fn.1.3.weight`,
`1.0.weight`, ...) and constructor arguments are identical, while `forward` runs the fused CUDA path
(engine.xit_forward / xit_backward) instead of ~25 ATen kernels.

The containers below only hold parameters; calling `XiT((x, y))` dispatches one autograd.Function.
"""
import torch
import torch.nn as nn

from . import engine, ops


class LayerNormBlock(nn.Module):
    """Holds ln_x / ln_y (finetune/xit.py:89-100)."""

    def __init__(self, emb_size):
        super().__init__()
        self.emb_size = emb_size
        self.ln_x = nn.LayerNorm(emb_size)
        self.ln_y = nn.LayerNorm(emb_size)


class MultiHeadAttention(nn.Module):
    """Holds the four projections (finetune/xit.py:113-124). `attention_mask='causal'` is accepted and,
    as in the reference (xit.py:134-140 discards the masked_fill result), has no effect."""

    def __init__(self, feat_size=768, num_heads=8, dropout=0, attention_mask="fully_visiable"):
        super().__init__()
        self.emb_size = feat_size
        self.num_heads = num_heads
        self.keys = nn.Linear(feat_size, feat_size)
        self.queries = nn.Linear(feat_size, feat_size)
        self.values = nn.Linear(feat_size, feat_size)
        self.att_drop = nn.Dropout(dropout)
        self.projection = nn.Linear(feat_size, feat_size)
        self.attention_mask = attention_mask
        if dropout != 0:
            raise ValueError("attention-probability dropout is 0 in every LR2PPO configuration; not implemented")


class FeedForwardBlock(nn.Sequential):
    def __init__(self, emb_size, expansion=4, drop_p=0.0):
        super().__init__(nn.Linear(emb_size, expansion * emb_size), nn.GELU(), nn.Dropout(drop_p),
                         nn.Linear(expansion * emb_size, emb_size))


class _Holder(nn.Module):
    """ResidualAdd / ResidualAddFusion: parameter containers named `fn` (finetune/xit.py:45-86)."""

    def __init__(self, fn):
        super().__init__()
        self.fn = fn


ResidualAdd = _Holder
ResidualAddFusion = _Holder


class XEncoderBlock(nn.Sequential):
    def __init__(self, feat_size=768, drop_p=0.1, forward_expansion=4, forward_drop_p=0.1, **kwargs):
        super().__init__(
            ResidualAddFusion(nn.Sequential(LayerNormBlock(feat_size), MultiHeadAttention(feat_size, **kwargs),
                                            nn.Dropout(drop_p))),
            ResidualAdd(nn.Sequential(nn.LayerNorm(feat_size),
                                      FeedForwardBlock(feat_size, expansion=forward_expansion, drop_p=forward_drop_p),
                                      nn.Dropout(drop_p))))


class XEncoder(nn.Sequential):
    def __init__(self, **kwargs):
        super().__init__(XEncoderBlock(**kwargs))


class XFeatureLayer(nn.Sequential):
    def __init__(self, feat_size=768):
        super().__init__(nn.LayerNorm(feat_size))


class _XiTFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, x, y, train, seed, save, *params):
        n, Sq, E = x.shape
        Skv = y.shape[1]
        W = engine.XitWeights(mod._bank, mod)
        xb = x.detach().reshape(n * Sq, E)
        yb = y.detach().reshape(n * Skv, E)
        xb = xb if xb.dtype == torch.bfloat16 else ops.to_bf16(xb.float().contiguous())
        yb = yb if yb.dtype == torch.bfloat16 else ops.to_bf16(yb.float().contiguous())
        out, c = engine.xit_forward(W, xb, yb, n, Sq, Skv, train, seed, 0, save)
        ctx.c, ctx.W, ctx.shapes, ctx.dt, ctx.nparams = c, W, (x.shape, y.shape), x.dtype, len(params)
        out = out.view(n, Sq, E)
        return out if x.dtype == torch.bfloat16 else ops.to_f32(out)

    @staticmethod
    def backward(ctx, dout):
        n, Sq, E = ctx.shapes[0]
        d = dout.reshape(n * Sq, E)
        d = d.contiguous() if d.dtype == torch.bfloat16 else ops.to_bf16(d.float().contiguous())
        sink = engine._GradSink()
        dx, dy = engine.xit_backward(ctx.W, ctx.c, d, sink)
        dx = dx.view(ctx.shapes[0]); dy = dy.view(ctx.shapes[1])
        if ctx.dt != torch.bfloat16:
            dx, dy = ops.to_f32(dx), ops.to_f32(dy)
        return (None, dx, dy, None, None, None) + (None,) * ctx.nparams


class XiT(nn.Sequential):
    """XiT(feat_size=768, **kwargs) called on a tuple (x [n,Sq,E], y [n,Skv,E]) -> [n,Sq,E]."""

    def __init__(self, feat_size=768, **kwargs):
        super().__init__(XEncoder(feat_size=feat_size, **kwargs), XFeatureLayer(feat_size=feat_size))
        self._bank = engine.ShadowBank()
        self._calls = 0

    def forward(self, x_y, **kwargs):
        x, y = x_y
        if not x.is_cuda:
            raise RuntimeError("lr2ppo_b200.XiT runs on CUDA (sm_100a) only; there is no CPU fallback")
        self._calls += 1
        seed = (torch.initial_seed() * 1000003 + self._calls) & 0x7FFFFFFFFFFFFFFF
        params = list(self.parameters())
        save = torch.is_grad_enabled() and (x.requires_grad or y.requires_grad or any(p.requires_grad for p in params))
        return _XiTFn.apply(self, x, y, self.training, seed, save, *params)
