"""TEST INFRASTRUCTURE ONLY — CPU fp32 restatement of ONE stage-3 LR2PPO step (rollout batch + update batch),
built from oracle/fusion_ref.py and oracle/restate.py.  Follows finetune/ppo.py:845-883 (rollout) and
:518-587 (update) with tencentpretrain AdamW (:374-402).  Used by tests (end-to-end parity) and by
bench.py's cpu_baseline / --impl reference legs (timed on the host cores).
Dropout is the identity here (the update in the reference runs in train mode with p = 0.1; parity tests
therefore drive the CUDA engine in eval mode, see tests/test_stage3_gpu.py)."""
import torch

from . import fusion_ref, restate

NO_DECAY = ("bias", "gamma", "beta")


class RefModel:
    """A state_dict of leaf tensors + AdamW state, for actor / critic / reward."""

    def __init__(self, sd, trainable=True):
        self.sd = {k: v.clone().requires_grad_(trainable) for k, v in sd.items()}
        self.m = {k: torch.zeros_like(v) for k, v in sd.items()} if trainable else None
        self.v = {k: torch.zeros_like(v) for k, v in sd.items()} if trainable else None

    def zero_grad(self):
        for p in self.sd.values():
            p.grad = None

    def adamw(self, lr):
        with torch.no_grad():
            for k, p in self.sd.items():
                if p.grad is None:
                    continue
                wd = 0.0 if any(nd in k for nd in NO_DECAY) else 0.01
                restate.adamw_step(p, p.grad, self.m[k], self.v[k], lr, wd)


def rollout(actor, critic, reward, text, img):
    """ref: finetune/ppo.py:845-883 with timestep 0: state = arange(tags)."""
    bs, T = text.shape[:2]
    with torch.no_grad():
        state = torch.arange(T, device=text.device).unsqueeze(0).repeat(bs, 1)
        scores = fusion_ref.actor_forward(actor.sd, text, img).view(bs, T)
        value = fusion_ref.critic_forward(critic.sd, text, img, state)
        _, idx = torch.sort(scores, dim=-1, descending=True, stable=True)
        next_state = torch.cat([torch.arange(2, device=text.device).unsqueeze(0).repeat(bs, 1),
                                torch.gather(state, 1, idx)], dim=1)
        rewards = fusion_ref.critic_forward(reward.sd, text, img, next_state)
    return state, next_state, scores, rewards, value


def update(actor, critic, memory, text, img, lr_actor, lr_critic, w_kl=0.001, w_ent=0.001, value_clip=0.5,
           actor_masks=None, critic_masks=None):
    """ref: finetune/ppo.py:518-587 (one stored batch).  actor_masks / critic_masks: the update runs in train mode in
    the reference (dropout 0.1 live, :515); pass the replayed masks (oracle/philox.py) to restate that, None = eval."""
    state, next_state, old_scores, rewards, old_value = memory
    bs, T = old_scores.shape
    actor.zero_grad(); critic.zero_grad()
    scores = fusion_ref.actor_forward(actor.sd, text, img, actor_masks).view(bs, T)
    value = fusion_ref.critic_forward(critic.sd, text, img, state, critic_masks)
    r = restate.ppo_policy_loss(scores, old_scores, rewards, old_value, next_state[:, -2:], w_kl, w_ent)
    r["loss"].backward()
    actor.adamw(lr_actor)
    vloss = restate.clipped_value_loss(value, r["reward_adj"].detach(), old_value, value_clip)
    vloss.backward()
    critic.adamw(lr_critic)
    return dict(policy_loss=r["loss"].detach(), value_loss=vloss.detach(), scores=scores.detach(),
                value=value.detach(), kl=r["kl"].detach(), entropy=r["entropy"].detach(), adv=r["adv"].detach(),
                rank_loss=r["rank_loss"].detach(), reward_adj=r["reward_adj"].detach())


def step(actor, critic, reward, text, img, lr_actor, lr_critic):
    mem = rollout(actor, critic, reward, text, img)
    return mem, update(actor, critic, mem, text, img, lr_actor, lr_critic)
