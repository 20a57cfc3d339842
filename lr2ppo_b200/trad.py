"""MSLR / MQ2008 ("trad") variants of the LR2PPO models (finetune/ppo_trad.py:142-283; the same classes appear
in ppo_eval_trad.py, reward_trad.py, pointwise_trad.py): each document is ONE 768-d row, XiT runs on a
single token (x, x), `out_layer` is Mlp(2*768 -> 3072 -> 768).  Built from the same fused pieces: the XiT
autograd Function, the Mlp Function and the row-dot head kernels.  `img_emb` is accepted and ignored, as in
the reference."""
import torch
import torch.nn as nn

from . import losses, ops
from .models import Mlp
from .xit import XiT


class _HeadFn(torch.autograd.Function):
    """logits[r] = <x[r*stride + off, :], w> + b  (768 -> 1 head, optionally on the last token of each group)."""

    @staticmethod
    def forward(ctx, x, w, b, rows, stride, off):
        xb = x.detach()
        xb = xb.contiguous() if xb.dtype == torch.bfloat16 else ops.to_bf16(xb.float().contiguous())
        xb = xb.view(-1, xb.shape[-1])
        ctx.save_for_backward(xb, w.detach())
        ctx.meta = (rows, stride, off, x.shape, x.dtype)
        return ops.rowdot_fwd(xb, w.detach().view(-1).contiguous(), b.detach().contiguous(), rows, stride, off)

    @staticmethod
    def backward(ctx, dout):
        xb, w = ctx.saved_tensors
        rows, stride, off, shape, dt = ctx.meta
        dx, dw, db = ops.rowdot_bwd(xb, w.view(-1).contiguous(), dout.contiguous().float(), rows, stride, off)
        dx = dx.view(shape)
        if dt != torch.bfloat16:
            dx = ops.to_f32(dx)
        return dx, dw.view_as(w), db, None, None, None


def linear_head(x, head, rows, stride=1, off=0):
    return _HeadFn.apply(x, head.weight, head.bias, rows, stride, off)


def _body(self, text_emb):
    """xit((x, x)) on single-token items, cat with the input, out_layer (ppo_trad.py:158-170)."""
    bs, tags = text_emb.shape[:2]
    x = text_emb.to(torch.float32).reshape(bs * tags, 1, 768)
    f = self.xit((x, x))
    f = torch.cat([f, x], dim=1)
    return self.out_layer(f.view(bs * tags, -1)), bs, tags


class Actor(nn.Module):
    def __init__(self, args, vit_args=None):
        super().__init__()
        self.mode = args.mode
        self.labels_num = args.labels_num
        self.xit = XiT(feat_size=768)
        self.out_layer = Mlp((1 + 1) * 768, 768 * 4, 768, nn.GELU, 0)
        if self.mode == "reg":
            self.head = nn.Linear(768, 1)
        else:
            raise ValueError("trad models are run with --mode reg in every shipped script")

    def forward(self, text_emb, img_emb, tgts):
        feat, bs, tags = _body(self, text_emb)
        logits = linear_head(feat, self.head, bs * tags).view(-1)
        if tgts is None:
            return logits
        return losses.smooth_l1_loss(logits.view(-1), tgts.view(-1), 0.3), logits


class Critic(nn.Module):
    def __init__(self, args, vit_args=None):
        super().__init__()
        self.mode = args.mode
        self.labels_num = args.labels_num
        self.pos_emb = nn.Embedding(4, 768)
        self.xit = XiT(feat_size=768)
        self.xitt = XiT(feat_size=768, attention_mask="causal")
        self.out_layer = Mlp((1 + 1) * 768, 768 * 4, 768, nn.GELU, 0)
        self.head = nn.Linear(768, 1)

    def forward(self, text_emb, img_emb, tgts, index):
        bs = text_emb.shape[0]
        bi = torch.arange(bs, device=text_emb.device).view(bs, 1)
        feat, bs, T = _body(self, text_emb[bi, index])
        x = feat.view(bs, T, 768) + self.pos_emb.weight[:T].unsqueeze(0)
        z = self.xitt((x, x))
        return linear_head(z, self.head, bs, T, T - 1).view(bs)


class Reward(Critic):
    pass


class ActorCritic(nn.Module):
    def __init__(self, args, vit_args=None):
        super().__init__()
        self.actor = Actor(args, vit_args)
        self.critic = Critic(args, vit_args)
