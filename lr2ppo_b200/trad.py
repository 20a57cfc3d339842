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


# ---- stage 3 on the trad models (BASELINE configs[0]: finetune/ppo_trad.py) ---------------------------------------
@torch.no_grad()
def rollout(model, reward_model, text_emb_batch, tgts_batch, state=None):
    """One rollout timestep (finetune/ppo_trad.py:765-813).  Returns the 7-entry memory
    [state, next_state, action_scores, rewards, value, text, tgts] that `train_model` consumes."""
    bs, tags_num = text_emb_batch.shape[:2]
    if state is None:
        state = torch.arange(tags_num, device=text_emb_batch.device).unsqueeze(0).repeat(bs, 1)
    was_training = model.training
    model.eval()
    reward_model.eval()
    action_scores = model.actor(text_emb_batch, None, None).view(bs, tags_num)
    value = model.critic(text_emb_batch, None, tgts_batch, state)
    next_state = ops.ppo_rollout(action_scores.contiguous(), state.contiguous(), 2)
    rewards = reward_model(text_emb_batch, None, tgts_batch, next_state)
    if was_training:
        model.train()
    return [state, next_state, action_scores, rewards, value, text_emb_batch, tgts_batch]


def train_model(args, model, optimizer, critic_optim, scheduler, critic_scheduler, memories, epoch):
    """ref: finetune/ppo_trad.py:432-544 — same arguments and the same ten returned averages; the per-row Python
    loop, the hinge-count sync and the ten per-batch all-reduces are replaced by the fused loss kernels."""
    total = None
    for state, next_state, old_action_prob, rewards, old_value, text, tgts in memories:
        model.zero_grad()
        bs, tags_num = old_action_prob.shape[:2]
        action_scores = model.actor(text, None, None).view(bs, tags_num)
        value = model.critic(text, None, tgts, state)
        pair = next_state[:, -2:].contiguous()
        loss, rank_loss, kl, ent, rewards_adj, adv = losses.ppo_policy_loss(
            action_scores, old_action_prob, rewards, old_value, pair, args.kl_div_loss_weight, args.entropy_weight,
            0.01, -0.1)
        loss.backward()
        optimizer.step()
        value_loss = losses.clipped_value_loss(value, rewards_adj.detach(), old_value, args.value_clip)
        value_loss.backward()
        critic_optim.step()
        stats = torch.stack([loss.detach(), value_loss.detach(), kl.mean(), old_value.mean(), value.detach().mean(),
                             rewards.mean(), rewards_adj.mean(), adv.mean(), rank_loss, ent.mean()])
        total = stats if total is None else total + stats
    total = total / len(memories)
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        total = total / dist.get_world_size()
        dist.all_reduce(total)
    scheduler.step()
    critic_scheduler.step()
    return list(total.unbind(0))
