"""Loss functions of the three training stages as autograd.Functions over the fused row kernels
(forward and analytic backward come out of the same launch)."""
import torch

from . import ops


class _PolicyLossFn(torch.autograd.Function):
    """ref: finetune/ppo.py:544-575 — KL penalty, entropy, advantage, pair order, RankLoss(0.01), policy loss."""

    @staticmethod
    def forward(ctx, s, s_old, reward, v_old, pi, w_kl, w_ent, margin, adv_eps):
        r = ops.ppo_policy_loss(s.detach().float().contiguous(), s_old.float().contiguous(),
                                reward.float().contiguous(), v_old.float().contiguous(), pi.contiguous(), w_kl, w_ent,
                                margin, adv_eps, want_grad=True)
        ctx.save_for_backward(r["ds"])
        ctx.mark_non_differentiable(r["kl"], r["entropy"], r["reward_adj"], r["adv"], r["rank_loss"])
        return r["loss"], r["rank_loss"], r["kl"], r["entropy"], r["reward_adj"], r["adv"]

    @staticmethod
    def backward(ctx, g, *unused):
        (ds,) = ctx.saved_tensors
        return ds * g, None, None, None, None, None, None, None, None


def ppo_policy_loss(scores, old_scores, rewards, old_values, next_state_pair, kl_weight, entropy_weight, margin=0.01,
                    adv_eps=-0.1):
    """Returns (loss, rank_loss, kl_penalty[B], entropy[B], rewards_adj[B], advantages[B]); gradient flows to scores."""
    return _PolicyLossFn.apply(scores, old_scores, rewards, old_values, next_state_pair, kl_weight, entropy_weight,
                               margin, adv_eps)


class _ValueLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, values, rewards, old_values, clip):
        loss, dv = ops.clipped_value_loss(values.detach().float().contiguous().view(-1),
                                          rewards.float().contiguous().view(-1),
                                          old_values.float().contiguous().view(-1), clip)
        ctx.save_for_backward(dv)
        ctx.shape = values.shape
        return loss

    @staticmethod
    def backward(ctx, g):
        (dv,) = ctx.saved_tensors
        return (dv * g).view(ctx.shape), None, None, None


def clipped_value_loss(values, rewards, old_values, clip):
    """ref: finetune/ppo.py:494-498 (same signature)."""
    return _ValueLossFn.apply(values, rewards, old_values, clip)


class _PairHingeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, chosen, reject, margin):
        loss, acc, dc, dr = ops.pair_hinge_loss(chosen.detach().float().contiguous(),
                                                reject.detach().float().contiguous(), margin)
        ctx.save_for_backward(dc, dr)
        ctx.mark_non_differentiable(acc)
        return loss, acc

    @staticmethod
    def backward(ctx, g, _):
        dc, dr = ctx.saved_tensors
        return dc * g, dr * g, None


def pair_hinge_loss(chosen, reject, margin=1.0):
    """ref: finetune/reward_pair_dataloader.py:355-358.  Returns (loss, accuracy)."""
    return _PairHingeFn.apply(chosen, reject, margin)


class _SmoothL1Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, tgt, beta):
        loss, dl = ops.smooth_l1_loss(logits.detach().float().contiguous().view(-1), tgt.contiguous().view(-1), beta)
        ctx.save_for_backward(dl)
        ctx.shape = logits.shape
        return loss

    @staticmethod
    def backward(ctx, g):
        (dl,) = ctx.saved_tensors
        return (dl * g).view(ctx.shape), None, None


def smooth_l1_loss(logits, tgts, beta=0.3):
    """ref: finetune/pointwise.py:229 / finetune/ppo.py:237 — nn.SmoothL1Loss(beta=0.3)(logits, int64 tgts)."""
    return _SmoothL1Fn.apply(logits, tgts, beta)


class RankLoss(torch.nn.Module):
    """RankLoss(margin)(scores, indices) with the reference's signature (finetune/ppo.py:38-55).
    Implemented with the policy-loss kernel with zero KL / entropy weights and unit |advantage|."""

    def __init__(self, margin=1):
        super().__init__()
        self.margin = margin

    def forward(self, scores, indices):
        B = scores.shape[0]
        one = torch.ones(B, device=scores.device)
        zero = torch.zeros(B, device=scores.device)
        # loss = rank_loss * mean|A| with A = 1 - 0 = 1  ->  equals the rank loss, gradients included
        loss, _, _, _, _, _ = ppo_policy_loss(scores, scores.detach(), one, zero, indices, 0.0, 0.0, self.margin, -0.1)
        return loss


# ---- north_star extensions (no counterpart in the reference's configuration: it ranks greedily and never reads
# --eps_clip; they complete the sampled-ranking PPO update the north_star names) ------------------------------------
class _RankLogProbFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, scores, perm):
        s = scores.detach().float().contiguous()
        ctx.save_for_backward(s, perm.contiguous())
        return ops.rank_logprob(s, perm.contiguous())

    @staticmethod
    def backward(ctx, g):
        s, perm = ctx.saved_tensors
        _, ds = ops.rank_logprob(s, perm, dlogprob=g.float().contiguous(), want_grad=True)
        return ds, None


def rank_logprob(scores, perm):
    """Plackett-Luce log-probability [B] of the rankings `perm` [B,n] (e.g. from ops.rank_sample) under `scores`
    [B,n]; differentiable in scores.  Evaluated on the scores that produced the sample it equals the sampler's own
    logprob bit for bit."""
    return _RankLogProbFn.apply(scores, perm)


class _ClipSurrogateFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logp, logp_old, adv, eps_clip, normalize, norm_eps):
        r = ops.ppo_clip_surrogate(logp.detach().float().contiguous().view(-1), logp_old.float().contiguous().view(-1),
                                   adv.float().contiguous().view(-1), eps_clip, normalize, norm_eps)
        ctx.save_for_backward(r["dlogp"])
        ctx.shape = logp.shape
        ctx.mark_non_differentiable(r["clip_frac"])
        return r["loss"], r["clip_frac"]

    @staticmethod
    def backward(ctx, g, _):
        (dl,) = ctx.saved_tensors
        return (dl * g).view(ctx.shape), None, None, None, None, None


def ppo_clip_surrogate(logp, logp_old, advantages, eps_clip=0.2, normalize=False, norm_eps=1e-5):
    """-mean(min(ratio * A, clamp(ratio, 1 - eps, 1 + eps) * A)), ratio = exp(logp - logp_old); returns
    (loss, clipped fraction).  normalize applies the reference's (unused) masked_normalize to the advantages
    (finetune/ppo.py:485-491).  Gradient flows to logp only."""
    return _ClipSurrogateFn.apply(logp, logp_old, advantages, eps_clip, normalize, norm_eps)
