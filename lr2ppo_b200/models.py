"""LR2PPO model classes with the reference's constructors, forward signatures and state_dict keys,
running on the fused sm_100a engine.

  Mlp                                 finetune/ppo.py:154-170
  Actor  (== stage-1 Classifier)      finetune/ppo.py:196-244, finetune/pointwise.py:189-236
  Critic / Reward (== stage-2 Classifier)   finetune/ppo.py:247-350, finetune/reward_pair_dataloader.py:233-283
  ActorCritic                         finetune/ppo.py:173-193

`args` needs: mode ('reg' | 'cls'), labels_num, seq_length, max_imgs, visual_feat_dim.
Inputs are fp32 CUDA tensors shaped as in the reference: text_emb [bs, tags, 196, 768],
img_emb [bs, tags, max_imgs, 768], tgts [bs, tags] int64, index [bs, 2|4] int64.
Outputs are fp32.  The compute dtype is bf16 with fp32 accumulation; master weights stay fp32.
"""
import torch
import torch.nn as nn

from . import engine, losses, ops
from .xit import XiT


class _MlpFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, x, save, *params):
        shp = x.shape
        xb = x.detach().reshape(-1, shp[-1])
        xb = xb.contiguous() if xb.dtype == torch.bfloat16 else ops.to_bf16(xb.float().contiguous())
        l1, l2 = engine._Lin(mod._bank, mod.fc1), engine._Lin(mod._bank, mod.fc2)
        y, c = engine.mlp_forward(l1, l2, xb, save)
        ctx.c, ctx.l, ctx.shp, ctx.dt, ctx.np, ctx.need_dx = c, (l1, l2), shp, x.dtype, len(params), x.requires_grad
        y = y.view(shp[:-1] + (y.shape[-1],))
        return y if x.dtype == torch.bfloat16 else ops.to_f32(y)

    @staticmethod
    def backward(ctx, dy):
        d = dy.reshape(-1, dy.shape[-1])
        d = d.contiguous() if d.dtype == torch.bfloat16 else ops.to_bf16(d.float().contiguous())
        dx = engine.mlp_backward(ctx.l[0], ctx.l[1], ctx.c, d, engine._GradSink(), need_dx=ctx.need_dx)
        if dx is not None:
            dx = dx.view(ctx.shp)
            dx = dx if ctx.dt == torch.bfloat16 else ops.to_f32(dx)
        return (None, dx, None) + (None,) * ctx.np


class Mlp(nn.Module):
    """fc1 -> exact GELU -> fc2 (dropout p as given; the reference always passes 0)."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.0):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        if act_layer is not nn.GELU or drop != 0.0:
            raise ValueError("lr2ppo_b200.Mlp implements the reference configuration: nn.GELU, drop=0")
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)
        self._bank = engine.ShadowBank()

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("lr2ppo_b200.Mlp runs on CUDA (sm_100a) only; there is no CPU fallback")
        params = list(self.parameters())
        save = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in params))
        return _MlpFn.apply(self, x, save, *params)


class _FusionFn(torch.autograd.Function):
    """One autograd node for a whole Actor / Critic / Reward forward; backward writes parameter
    gradients straight into `.grad` (fp32), like AccumulateGrad would."""

    @staticmethod
    def forward(ctx, eng, text, img, index, train, save, *params):
        logits, c = eng.forward(text, img, index, train=train, save=save)
        ctx.eng, ctx.c, ctx.np = eng, c, len(params)
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        ctx.eng.backward(ctx.c, dlogits)
        ctx.c = None
        return (None, None, None, None, None, None) + (None,) * ctx.np


def _check_inputs(text_emb, img_emb):
    if not (text_emb.is_cuda and img_emb.is_cuda):
        raise RuntimeError("lr2ppo_b200 models run on CUDA (sm_100a) only; there is no CPU fallback")
    ok = (torch.float32, torch.bfloat16)
    if text_emb.dtype not in ok or img_emb.dtype not in ok:
        raise RuntimeError("text_emb / img_emb must be fp32 (as read from clean_feat.h5) or bf16 (stored rollout "
                           "batches, ppo.RolloutMemory)")


class Actor(nn.Module):
    def __init__(self, args, vit_args=None):
        super().__init__()
        self.mode = args.mode
        self.labels_num = args.labels_num
        self.text_proj = Mlp(768, 768 * 4, 768, nn.GELU, 0)
        self.img_proj = Mlp(768, 768 * 4, 768, nn.GELU, 0)
        self.xit = XiT(feat_size=768)
        self.out_layer = Mlp((args.seq_length + args.max_imgs) * args.visual_feat_dim, 768 * 4, 768, nn.GELU, 0)
        if self.mode == "cls":
            self.head = nn.Linear(768, self.labels_num)
        elif self.mode == "reg":
            self.head = nn.Linear(768, 1)
        self._engine = engine.FusionEngine(self, "actor")
        self._logit_shape = (-1,)

    def scores(self, text_emb, img_emb):
        """logits of every (clip, tag) item: [bs*tags] ('reg') or [bs*tags, labels_num] ('cls')."""
        _check_inputs(text_emb, img_emb)
        params = list(self.parameters())
        save = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        logits = _FusionFn.apply(self._engine, text_emb.contiguous(), img_emb.contiguous(), None, self.training, save,
                                 *params)
        return logits

    def forward(self, text_emb, img_emb, tgts):
        logits = self.scores(text_emb, img_emb)
        if self.mode == "cls":
            logits = logits.view(-1, self.labels_num)
        else:
            logits = logits.view(*self._logit_shape)
        if self.mode == "reg":
            if tgts is None:
                return logits
            return losses.smooth_l1_loss(logits.view(-1), tgts.view(-1), 0.3), logits
        if tgts is not None:
            loss = nn.NLLLoss()(nn.LogSoftmax(dim=-1)(logits), tgts.view(-1))
            return loss, logits
        return logits


class Classifier(Actor):
    """Stage-1 pointwise model (finetune/pointwise.py:189-236): Actor with logits shaped [bs*tags, 1]."""

    def __init__(self, args, vit_args=None):
        super().__init__(args, vit_args)
        self._logit_shape = (-1, 1)


class Critic(nn.Module):
    def __init__(self, args, vit_args=None):
        super().__init__()
        self.mode = args.mode
        self.labels_num = args.labels_num
        self.text_proj = Mlp(768, 768 * 4, 768, nn.GELU, 0)
        self.img_proj = Mlp(768, 768 * 4, 768, nn.GELU, 0)
        self.pos_emb = nn.Embedding(4, 768)
        self.xit = XiT(feat_size=768)
        self.xitt = XiT(feat_size=768, attention_mask="causal")
        self.out_layer = Mlp((args.seq_length + args.max_imgs) * args.visual_feat_dim, 768 * 4, 768, nn.GELU, 0)
        self.head = nn.Linear(768, 1)
        self._engine = engine.FusionEngine(self, "critic")

    def forward(self, text_emb, img_emb, tgts, index):
        _check_inputs(text_emb, img_emb)
        if index.shape[1] > self.pos_emb.weight.shape[0]:
            raise RuntimeError("index longer than pos_emb")
        params = list(self.parameters())
        save = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        return _FusionFn.apply(self._engine, text_emb.contiguous(), img_emb.contiguous(),
                               index.to(torch.int64).contiguous(), self.training, save, *params)


class Reward(Critic):
    """Same architecture and forward as Critic; index is the 4-slot [0, 1, pi(0), pi(1)] layout
    (finetune/ppo.py:300-350).

    The frozen reward model scores a permutation of items whose features do not depend on the permutation, so its
    inference forward can be cut in two (ppo.rollout runs the first half beside the actor, on its own stream):
    item_features(text, img) -> pooled features of every (clip, tag) item; from_item_features(feat, index) -> rewards.
    from_item_features(item_features(t, i), idx) == forward(t, i, tgts, idx) bit for bit in eval mode."""

    @torch.no_grad()
    def item_features(self, text_emb, img_emb):
        _check_inputs(text_emb, img_emb)
        if self.training:
            raise RuntimeError("item_features is inference-only (model.eval())")
        return self._engine.forward(text_emb.contiguous(), img_emb.contiguous(), None, train=False, save=False,
                                    trunk_only=True)[0]

    @torch.no_grad()
    def from_item_features(self, feat, index):
        if index.shape[1] > self.pos_emb.weight.shape[0]:
            raise RuntimeError("index longer than pos_emb")
        return self._engine.tail(feat, index.to(torch.int64).contiguous())


class PairClassifier(Critic):
    """Stage-2 reward model `Classifier` (finetune/reward_pair_dataloader.py:233-283)."""


class ActorCritic(nn.Module):
    def __init__(self, args, vit_args=None):
        super().__init__()
        self.actor = Actor(args, vit_args)
        self.critic = Critic(args, vit_args)

    def enable_actor(self):
        for p in self.actor.parameters():
            p.requires_grad = True

    def disable_actor(self):
        for p in self.actor.parameters():
            p.requires_grad = False

    def enable_critic(self):
        for p in self.critic.parameters():
            p.requires_grad = True

    def disable_critic(self):
        for p in self.critic.parameters():
            p.requires_grad = False
