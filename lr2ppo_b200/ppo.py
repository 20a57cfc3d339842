"""Stage-3 (LR2PPO) rollout, update and evaluation with the reference's function signatures
(finetune/ppo.py: build_optimizer:378, clipped_value_loss:494, train_model:501, evaluate:620, rollout
loop :845-883), re-built on the fused engine:

  * rollout: actor + critic + stable descending sort + permutation compose + reward model, no host sync
    (the reference loops over the batch in Python, finetune/ppo.py:869-871);
  * update: one fused kernel for KL / entropy / advantage / pair order / RankLoss / policy loss and its
    backward (the reference synchronises the host once per row at :563-568 and again at :52, :576);
  * the ten per-batch statistic all-reduces (:589-598) are packed into one.
"""
import os

import torch
import torch.distributed as dist

from . import ops
from .losses import RankLoss, clipped_value_loss, ppo_policy_loss  # noqa: F401  (re-exported, reference names)
from .models import Actor, ActorCritic, Critic, Mlp, Reward  # noqa: F401
from .ndcg import AverageNDCGMeter
from .optim import attach_shadows, decay_groups, make_scheduler, str2optimizer, str2scheduler  # noqa: F401


_SIDE = {}


def _reward_stream(device):
    """Third stream of the rollout (only with the two-branch step on): the reward model's item features."""
    if _branch_stream(device) is None or os.environ.get("LR2_REWARD_STREAM", "1") != "1":
        return None
    key = ("reward", device.index if device.index is not None else torch.cuda.current_device())
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device=device)
    return _SIDE[key]


def _branch_stream(device):
    """Second CUDA stream for the critic's branch of a step: the actor and the critic are independent models whose
    only coupling inside a step is one [bs] vector (the adjusted rewards the value loss regresses on), so their
    forwards / backwards / optimizer passes run as two branches of the step's CUDA graph; the launch-latency-bound
    kernels of one (LayerNorm, bias column sums, 2-4-token attention, tiny GEMMs) and, data parallel, its collectives
    fill the gaps of the other.  Measured on B200: 10.08 -> 9.78 ms/step at N = 1, 8.39 -> 7.89 ms at N = 2, results
    bit-identical.  LR2_DUAL_STREAM=0 (-> None) runs the step on one stream, in the reference's order."""
    if os.environ.get("LR2_DUAL_STREAM", "1") != "1" or device.type != "cuda":
        return None
    key = device.index if device.index is not None else torch.cuda.current_device()
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device=device)
    return _SIDE[key]


def log(t, eps=1e-20):
    """ref: finetune/ppo.py:431-432."""
    return torch.log(t.clamp(min=eps))


def build_optimizer(args, model):
    """ref: finetune/ppo.py:378-419 — two AdamW optimizers (actor, critic), no decay for bias/gamma/beta,
    linear warm-up schedules.  Returns (optimizer, critic_optimizer, scheduler, critic_scheduler)."""
    opt_name = getattr(args, "optimizer", "adamw")
    if opt_name not in str2optimizer:
        raise ValueError(f"optimizer {opt_name!r} is outside the LR2PPO hot path (only adamw is used by the scripts)")
    optimizer = str2optimizer[opt_name](decay_groups(model.actor.named_parameters()), lr=args.learning_rate,
                                        correct_bias=False)
    critic_optimizer = str2optimizer[opt_name](decay_groups(model.critic.named_parameters()),
                                               lr=args.critic_learning_rate, correct_bias=False)
    for eng, opt in ((getattr(model.actor, "_engine", None), optimizer),
                     (getattr(model.critic, "_engine", None), critic_optimizer)):
        if eng is None:          # trad (MSLR) models: plain autograd Functions; FusedAdamW bumps the parameters'
            continue             # version counters, so their ShadowBank copies are re-cast after every step
        attach_shadows(eng, opt)
        # opt-in: correct, but its drain keeps too few loads in flight today (3.2 TB/s), so it is slower
        # than wgrad + AdamW; LR2_WGRAD_ADAMW_IMPL=mma selects the linear-pass variant (adamw_wgrad.cu)
        if getattr(args, "fused_fc1", False):
            eng.enable_fused_fc1(opt)
        elif getattr(args, "fc1_grad_bf16", False):
            eng.enable_bf16_fc1_grad(opt)
    return optimizer, critic_optimizer, make_scheduler(args, optimizer), make_scheduler(args, critic_optimizer)


@torch.no_grad()
def rollout(model, reward_model, text_emb_batch, img_emb_batch, tgts_batch, state=None, before_critic=None):
    """One rollout timestep (finetune/ppo.py:845-883).  Returns the memory entry
    [state, next_state, action_scores, rewards, value, text, img, tgts] (no clones needed: nothing aliases).
    The critic's value does not depend on the actor / reward results, so it is evaluated last; `before_critic()` (the
    pending all-gather of the critic's weights in the pipelined data-parallel step) is called right before it."""
    bs, tags_num = text_emb_batch.shape[:2]
    if state is None:
        state = torch.arange(tags_num, device=text_emb_batch.device).unsqueeze(0).repeat(bs, 1)
    was_training = model.training
    model.eval()
    reward_model.eval()
    side = _branch_stream(text_emb_batch.device)
    if side is not None:                # the value needs neither the actor's nor the reward model's result
        main = torch.cuda.current_stream()
        side.wait_stream(main)
        with torch.cuda.stream(side):
            if before_critic is not None:
                before_critic()
            value = model.critic(text_emb_batch, img_emb_batch, tgts_batch, state)
    # the reward model scores a permutation of items whose pooled features do not depend on it: that part (all of the
    # forward up to out_layer) runs on a third stream beside the actor; only gather + xitt + head wait for next_state
    rstream = _reward_stream(text_emb_batch.device) if hasattr(reward_model, "item_features") else None
    if rstream is not None:
        rstream.wait_stream(main)
        with torch.cuda.stream(rstream):
            item_feat = reward_model.item_features(text_emb_batch, img_emb_batch)
    action_logits = model.actor.scores(text_emb_batch, img_emb_batch)
    if model.actor.mode == "cls":
        pr = action_logits.view(bs, tags_num, 3).softmax(dim=-1)
        action_scores = pr[:, :, 1] + 2 * pr[:, :, 2]
    else:
        action_scores = action_logits.view(bs, tags_num)
    # timestep > 0 passes the previous next_state (prefix + permutation) as state: the reference's index_select with
    # the tags_num sort indices reads its first tags_num entries (finetune/ppo.py:869-871)
    next_state = ops.ppo_rollout(action_scores.contiguous(), state[:, :tags_num].contiguous(), 2)
    if rstream is not None:
        main.wait_stream(rstream)
        rewards = reward_model.from_item_features(item_feat, next_state)
    else:
        rewards = reward_model(text_emb_batch, img_emb_batch, tgts_batch, next_state)
    if side is not None:
        main.wait_stream(side)
    else:
        if before_critic is not None:
            before_critic()
        value = model.critic(text_emb_batch, img_emb_batch, tgts_batch, state)
    if was_training:
        model.train()
    return [state, next_state, action_scores, rewards, value, text_emb_batch, img_emb_batch, tgts_batch]


def _sync_and_step(grad_sync, module, optimizer):
    """optimizer.step() (finetune/ppo.py:580,587); data-parallel: the all-reduce of the small gradients runs on NCCL's
    stream while AdamW already updates out_layer.fc1 (97 % of the parameters; its gradient was built from all-gathered
    operands and needs no reduction).  Returns wait(): with a row-sharded fc1 it completes the all-gather of the
    updated bf16 shadow rows and must be called before the module's next forward."""
    if grad_sync is None:
        optimizer.step()
        return lambda: None
    early = grad_sync.early_params(module)
    if early and hasattr(optimizer, "register_shadow"):
        optimizer.step(first=early, between=grad_sync.start(module))
    else:
        grad_sync(module)
        optimizer.step()
    return grad_sync.after_step(module)


_STAT_NAMES = ["policy_loss", "value_loss", "kl_penalty", "old_value", "value", "rewards_ori", "rewards",
               "advantages", "rank_loss", "entropy"]


def update_batch(args, model, optimizer, critic_optim, memory, grad_sync=None, defer_critic_wait=False):
    """One stored batch of the update loop (finetune/ppo.py:518-587).  Returns a [10] tensor of the
    statistics the reference logs (means over the batch), still on the device.  defer_critic_wait: return
    (stats, wait) instead, where wait() completes the critic's pending weight all-gather (row-sharded data parallel)
    and must be called before the critic's next forward."""
    state, next_state, old_action_prob, rewards, old_value, text, img, tgts = memory
    if model.actor._engine.persistent_grads:
        model.actor._engine.begin_step(); model.critic._engine.begin_step()   # .grad buffers are reused
    else:
        model.zero_grad()
    bs, tags_num = old_action_prob.shape[:2]
    side = _branch_stream(text.device)
    if side is not None:
        return _update_batch_two_branches(args, model, optimizer, critic_optim, memory, side, grad_sync,
                                          defer_critic_wait)
    action_logits = model.actor.scores(text, img)
    value = model.critic(text, img, tgts, state)
    if model.actor.mode == "cls":
        pr = action_logits.view(bs, tags_num, 3).softmax(dim=-1)
        action_scores = pr[:, :, 1] + 2 * pr[:, :, 2]
    else:
        action_scores = action_logits.view(bs, tags_num)
    pair = next_state[:, -2:].contiguous()
    loss, rank_loss, kl, ent, rewards_adj, adv = ppo_policy_loss(
        action_scores, old_action_prob, rewards, old_value, pair, args.kl_div_loss_weight, args.entropy_weight,
        0.01, -0.1)
    loss.backward()
    wait_actor = _sync_and_step(grad_sync, model.actor, optimizer)       # shadow gather overlaps the critic's work
    value_loss = clipped_value_loss(value, rewards_adj.detach(), old_value, args.value_clip)
    value_loss.backward()
    wait_critic = _sync_and_step(grad_sync, model.critic, critic_optim)
    wait_actor()
    stats = torch.stack([loss.detach(), value_loss.detach(), kl.mean(), old_value.mean(), value.detach().mean(),
                         rewards.mean(), rewards_adj.mean(), adv.mean(), rank_loss, ent.mean()])
    if defer_critic_wait:
        return stats, wait_critic
    wait_critic()
    return stats


def _update_batch_two_branches(args, model, optimizer, critic_optim, memory, side, grad_sync, defer_critic_wait):
    """update_batch with the critic on its own stream: critic forward || actor forward; the value loss waits for the
    policy-loss kernel (its regression target); critic backward + AdamW || actor backward + AdamW.  Same kernels,
    same arithmetic, same results as the single-stream order.  Data parallel: every rank issues the collectives of the
    two branches in the same program order, so they pair up across ranks exactly as before, and the communication of
    one model overlaps the computation of the other."""
    state, next_state, old_action_prob, rewards, old_value, text, img, tgts = memory
    bs, tags_num = old_action_prob.shape[:2]
    main = torch.cuda.current_stream()
    side.wait_stream(main)
    with torch.cuda.stream(side):
        value = model.critic(text, img, tgts, state)
    action_logits = model.actor.scores(text, img)
    if model.actor.mode == "cls":
        pr = action_logits.view(bs, tags_num, 3).softmax(dim=-1)
        action_scores = pr[:, :, 1] + 2 * pr[:, :, 2]
    else:
        action_scores = action_logits.view(bs, tags_num)
    pair = next_state[:, -2:].contiguous()
    loss, rank_loss, kl, ent, rewards_adj, adv = ppo_policy_loss(
        action_scores, old_action_prob, rewards, old_value, pair, args.kl_div_loss_weight, args.entropy_weight,
        0.01, -0.1)
    target = rewards_adj.detach()
    ready = torch.cuda.Event()
    ready.record(main)
    with torch.cuda.stream(side):
        side.wait_event(ready)
        value_loss = clipped_value_loss(value, target, old_value, args.value_clip)
        value_loss.backward()
        wait_critic = _sync_and_step(grad_sync, model.critic, critic_optim)
    loss.backward()
    wait_actor = _sync_and_step(grad_sync, model.actor, optimizer)
    wait_actor()
    if not defer_critic_wait:
        with torch.cuda.stream(side):
            wait_critic()
    main.wait_stream(side)
    stats = torch.stack([loss.detach(), value_loss.detach(), kl.mean(), old_value.mean(), value.detach().mean(),
                         rewards.mean(), rewards_adj.mean(), adv.mean(), rank_loss, ent.mean()])
    if defer_critic_wait:
        return stats, wait_critic
    return stats


def train_model(args, model, optimizer, critic_optim, scheduler, critic_scheduler, memories, epoch, grad_sync=None):
    """ref: finetune/ppo.py:501-617 — same arguments, same ten returned averages
    [policy_loss, value_loss, kl_penalty, old_value, value, rewards_ori, rewards, advantages, rank_loss, entropy]."""
    total = None
    for memory in memories:
        stats = update_batch(args, model, optimizer, critic_optim, memory, grad_sync)
        total = stats if total is None else total + stats
    total = total / len(memories)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        total = total / dist.get_world_size()
        dist.all_reduce(total)          # one packed all-reduce instead of ten per batch
    scheduler.step()
    critic_scheduler.step()
    return list(total.unbind(0))


@torch.no_grad()
def evaluate_scores(model, text_emb, img_emb):
    """Actor scores of every tag of one clip (finetune/ppo.py:640-649)."""
    logits = model.actor.scores(text_emb, img_emb)
    if model.actor.mode == "cls":
        logits = logits.view(-1, 3)
        return logits[:, 1] + 2 * logits[:, 2]
    return logits.view(-1)


@torch.no_grad()
def evaluate(args, val_loader, step, split="test", num_tasks=None, cases_path=None):
    """ref: finetune/ppo.py:620-681 and finetune/ppo_eval.py:401-471.  Scores every clip of this rank's shard, then ONE
    segmented NDCG launch and ONE all_gather for the whole pass (the reference gathers once per clip).

    The loader yields (text_emb [1, tags, S, E], img_emb [1, I, E], tgts [1, tags]) or, for ppo_eval, a fourth element:
    the clip record as default_collate batches it.  cases_path: write the per-clip dump of ppo_eval.py:441-459
    (`case/ppo_cases.json`: filename, id, description, tags, ndcg[6], predict = tags in predicted order with scores)."""
    args.model.eval()
    scores_l, gold_l, clips = [], [], []
    # The reference scores one clip per forward (batch_size 1, finetune/ppo.py:697-698), i.e. it streams the whole
    # 500 M-parameter out_layer weight once per clip.  Tags of different clips are independent for the actor, so the
    # (clip, tag) items of consecutive clips are packed into one forward of up to `eval_items` items.
    cap = int(getattr(args, "eval_items", 192))
    pend_t, pend_i, pend_n = [], [], []

    def flush():
        if not pend_t:
            return
        text = torch.cat(pend_t, dim=0)                     # [items, 1, S, E]
        img = torch.cat(pend_i, dim=0)                      # [items, 1, I, E]
        sc = evaluate_scores(args.model, text, img)
        off = 0
        for n_tags in pend_n:
            scores_l.append(sc[off:off + n_tags])
            off += n_tags
        pend_t.clear(); pend_i.clear(); pend_n.clear()

    for batch in val_loader:
        text_emb, img_emb, tgts = batch[:3]
        if len(batch) > 3:
            clips.append(batch[3])
        n_tags = text_emb.shape[1]
        if pend_n and sum(pend_n) + n_tags > cap:
            flush()
        text = text_emb.to(args.device, non_blocking=True)  # [1, tags, S, E]
        pend_t.append(text.view(n_tags, 1, *text.shape[2:]))
        pend_i.append(img_emb.to(args.device, non_blocking=True).unsqueeze(1)
                      .expand(1, n_tags, *img_emb.shape[1:]).reshape(n_tags, 1, *img_emb.shape[1:]))
        pend_n.append(n_tags)
        gold_l.append(tgts.to(args.device, non_blocking=True).view(-1))
    flush()
    meter = AverageNDCGMeter()
    n = len(scores_l)
    nmax = max(s.numel() for s in scores_l)
    scores = torch.full((n, nmax), float("-inf"), device=args.device)
    labels = torch.zeros((n, nmax), dtype=torch.int64, device=args.device)
    lens = torch.tensor([s.numel() for s in scores_l], dtype=torch.int32, device=args.device)
    for i, (s, g) in enumerate(zip(scores_l, gold_l)):
        scores[i, :s.numel()] = s
        labels[i, :g.numel()] = g
    want_cases = cases_path is not None and clips
    res = meter.batch_ndcg(scores, labels, lens=lens, want_order=bool(want_cases))
    vals, order = res if want_cases else (res, None)
    if want_cases:
        _dump_cases(cases_path, clips, vals.cpu(), order.cpu(), scores.cpu())
    if num_tasks and num_tasks > 1:
        gathered = [torch.zeros_like(vals) for _ in range(num_tasks)]
        dist.all_gather(gathered, vals)
        vals = torch.cat(gathered, dim=0)
    if getattr(args, "is_master", True):
        meter.add_batch(vals)
        ndcg_value = meter.value()
        if hasattr(args, "logger"):
            args.logger.info("NDCG:")
            args.logger.info("".join("\nNDCG@{}={:.4f}".format(k, ndcg_value[k]) for k in sorted(ndcg_value.keys())))
        return ndcg_value[100000000]
    return None


def _plain(v):
    """A collated clip field -> what json.dump writes for it in the reference (lists of strings stay lists)."""
    if torch.is_tensor(v):
        return v.cpu().tolist()
    return v


def _dump_cases(path, clips, ndcg, order, scores):
    """ppo_eval.py:441-459: one record per clip.  `clip` is the default_collate'd record (batch size 1): strings
    arrive as 1-element lists, each tag's target as a 1-element tensor."""
    import json
    import os
    results = []
    for i, clip in enumerate(clips):
        rec = {key: _plain(clip[key]) for key in ("filename", "id", "description")}
        rec["tags"] = [{"tag": _plain(t["tag"]), "target": int(torch.as_tensor(t["target"]).view(-1)[0])}
                       for t in clip["tags"]]
        rec["ndcg"] = ndcg[i].tolist()
        n = len(rec["tags"])
        rec["predict"] = [(rec["tags"][j], float(scores[i, j])) for j in order[i, :n].tolist()]
        results.append(rec)
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    with open(path, "w") as f:
        json.dump(results, f)


class GraphedStage3Step:
    """One stage-3 step (rollout + update of one batch) captured in a CUDA graph and replayed.

    ~300 kernel launches + the torch glue of a step cost more host time than the small kernels take on the
    GPU; replaying a graph removes that.  Requirements handled here: static input buffers, persistent `.grad`
    buffers (no allocation / pointer-table rebuild inside the graph), AdamW hyper-parameters refreshed from
    the host before each replay, dropout seeds read from a device counter bumped inside the graph.
    With `grad_sync` (N > 1) the NCCL all-gather / all-reduce calls are captured in the same graph."""

    def __init__(self, args, model, reward_model, optimizer, critic_optim, text, img, tgts, warmup=3, grad_sync=None):
        self.args, self.model, self.reward, self.grad_sync = args, model, reward_model, grad_sync
        self.opt, self.copt = optimizer, critic_optim
        self.text, self.img, self.tgts = text.clone(), img.clone(), tgts.clone()
        for e in (model.actor._engine, model.critic._engine):
            e.persistent_grads = True
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.stats = self._eager()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.opt.frozen_hyper = self.copt.frozen_hyper = True
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.stats = self._eager()

    def _eager(self):
        mem = rollout(self.model, self.reward, self.text, self.img, self.tgts)
        self.model.train()
        stats = update_batch(self.args, self.model, self.opt, self.copt, mem, self.grad_sync)
        self.model.eval()
        return stats

    def static_inputs(self):
        """The graph's input buffers (text, img, tgts); fill them (e.g. feed.DeviceFeeder.next_into) then replay()."""
        return self.text, self.img, self.tgts

    def replay(self):
        self.opt.update_hyper(); self.copt.update_hyper()
        self.graph.replay()
        return self.stats

    def __call__(self, text, img, tgts):
        self.text.copy_(text, non_blocking=True)
        self.img.copy_(img, non_blocking=True)
        self.tgts.copy_(tgts, non_blocking=True)
        return self.replay()


class PipelinedStage3Step(GraphedStage3Step):
    """Same training loop, cut at a different point: one replay = update of the PREVIOUS batch followed by the rollout
    of the current one (the sequence rollout 0, update 0, rollout 1, update 1, ... is unchanged, so are all results).
    With the row-sharded data-parallel optimizer this lets the all-gather of the critic's updated weights -- the last
    thing an update does -- run under the next rollout's actor and reward forwards inside the same CUDA graph instead
    of being exposed at the end of the step.  `replay()` returns the statistics of the update it contained (i.e. of
    the batch fed one call earlier); `flush()` runs the final pending update eagerly."""

    def __init__(self, args, model, reward_model, optimizer, critic_optim, text, img, tgts, warmup=2, grad_sync=None):
        self.args, self.model, self.reward, self.grad_sync = args, model, reward_model, grad_sync
        self.opt, self.copt = optimizer, critic_optim
        self.text, self.img, self.tgts = text.clone(), img.clone(), tgts.clone()
        for e in (model.actor._engine, model.critic._engine):
            e.persistent_grads = True
        mem = rollout(model, reward_model, self.text, self.img, self.tgts)             # prologue: rollout of batch 0
        self.mem = [t.clone() for t in mem]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.stats = self._eager()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.opt.frozen_hyper = self.copt.frozen_hyper = True
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.stats = self._eager()

    def _eager(self):
        self.model.train()
        stats, wait_critic = update_batch(self.args, self.model, self.opt, self.copt, self.mem, self.grad_sync,
                                          defer_critic_wait=True)
        self.model.eval()
        new = rollout(self.model, self.reward, self.text, self.img, self.tgts, before_critic=wait_critic)
        for dst, src in zip(self.mem, new):
            dst.copy_(src)
        return stats

    def flush(self):
        """Run the update of the last rolled-out batch (eagerly) and return its statistics."""
        frozen = self.opt.frozen_hyper
        self.opt.frozen_hyper = self.copt.frozen_hyper = False
        self.model.train()
        stats = update_batch(self.args, self.model, self.opt, self.copt, self.mem, self.grad_sync)
        self.model.eval()
        self.opt.frozen_hyper = self.copt.frozen_hyper = frozen
        return stats


# ---------------------------------------------------------------------------------------------------------------------
# The reference's real loop shape (finetune/ppo.py:827-908): `update_timesteps` (200) rollout batches are collected,
# each keeping CLONES of its fp32 text / image features in a Python list (:882-883, 31 MB per entry, 6.3 GB per
# cycle), then train_model walks the list.  SURVEY.md §8(f) 1: a preallocated bf16 ring buffer instead.
# ---------------------------------------------------------------------------------------------------------------------
class RolloutMemory:
    """Preallocated device ring of `capacity` rollout batches.  text / img are stored ONCE in bf16 -- the precision
    every kernel of the path consumes them in, so nothing changes numerically: the engine's gather + cast of the
    first forward simply happens here -- and all later forwards (the rollout's three models, the update's two) read
    the slot in place.  img is stored un-repeated ([bs, 1, I, E]).  Per 24-query batch: 15.1 MB instead of the
    reference's 31.3 MB of fp32 clones; a 200-batch cycle holds 3.0 GB instead of 6.3 GB.

        mem = RolloutMemory(200, 24, 2, 196, 16, 768, device)
        slot = mem.stage(text_fp32, img_fp32, tgts)          # cast into the next free slot, returns its views
        entry = rollout(model, reward_model, *slot)          # runs on the bf16 views
        mem.commit(entry)                                    # small tensors (state, scores, rewards, value) copied in
        for entry in mem: update_batch(...)                  # entries in insertion order
        mem.clear()
    """

    def __init__(self, capacity, bs, tags, S, I, E, device, n_prefix=2):
        self.capacity, self.n = int(capacity), 0
        kw = dict(device=device)
        self.text = torch.empty((capacity, bs, tags, S, E), dtype=torch.bfloat16, **kw)
        self.img = torch.empty((capacity, bs, 1, I, E), dtype=torch.bfloat16, **kw)
        self.tgts = torch.zeros((capacity, bs, tags), dtype=torch.int64, **kw)
        self.state = torch.zeros((capacity, bs, tags), dtype=torch.int64, **kw)
        self.next_state = torch.zeros((capacity, bs, n_prefix + tags), dtype=torch.int64, **kw)
        self.scores = torch.zeros((capacity, bs, tags), dtype=torch.float32, **kw)
        self.rewards = torch.zeros((capacity, bs), dtype=torch.float32, **kw)
        self.value = torch.zeros((capacity, bs), dtype=torch.float32, **kw)

    def __len__(self):
        return self.n

    def bytes_per_entry(self):
        return sum(t[0].numel() * t.element_size() for t in (self.text, self.img, self.tgts, self.state,
                                                             self.next_state, self.scores, self.rewards, self.value))

    def stage(self, text, img, tgts):
        """Cast one fp32 (or copy one bf16) batch into the next slot; returns (text, img, tgts) views of the slot.
        img: [bs, I, E] as the loader yields it, or [bs, 1, I, E]."""
        if self.n >= self.capacity:
            raise RuntimeError("RolloutMemory is full: run the update and clear() first")
        k = self.n
        img = img.view(img.shape[0], 1, *img.shape[-2:]) if img.dim() == 3 else img[:, :1]
        for dst, src in ((self.text[k], text), (self.img[k], img)):
            if src.dtype == torch.float32:
                ops.to_bf16(src.contiguous(), out=dst)
            else:
                dst.copy_(src)
        self.tgts[k].copy_(tgts)
        return self.text[k], self.img[k], self.tgts[k]

    def commit(self, entry):
        """entry: the list rollout() returned for the staged slot."""
        k = self.n
        state, next_state, scores, rewards, value = entry[:5]
        self.state[k].copy_(state); self.next_state[k].copy_(next_state); self.scores[k].copy_(scores)
        self.rewards[k].copy_(rewards.view(-1)); self.value[k].copy_(value.view(-1))
        self.n += 1

    def entry(self, k):
        return [self.state[k], self.next_state[k], self.scores[k], self.rewards[k], self.value[k], self.text[k],
                self.img[k], self.tgts[k]]

    def __iter__(self):
        return (self.entry(k) for k in range(self.n))

    def clear(self):
        self.n = 0


class GraphedCycle:
    """The reference's cycle -- N rollout batches, then N update batches over the stored memory, schedulers stepped
    once per cycle (finetune/ppo.py:845-908, 612-613) -- with both inner loops replayed from CUDA graphs:

      * ONE rollout graph reading a static staging slot (the batch is cast into it, the graph writes its results into
        static small tensors, both are then copied into the ring slot: 15 MB device-to-device per batch);
      * ONE update graph reading a static "current entry" (an entry is copied in, the graph replays).

    Statistics of the cycle are accumulated on the device and returned as the reference's ten averages."""

    def __init__(self, args, model, reward_model, optimizer, critic_optim, capacity, bs, tags, S=196, I=16, E=768,
                 grad_sync=None, warmup=2):
        dev = next(model.parameters()).device
        self.args, self.model, self.reward, self.grad_sync = args, model, reward_model, grad_sync
        self.opt, self.copt = optimizer, critic_optim
        self.memory = RolloutMemory(capacity, bs, tags, S, I, E, dev)
        self.cur = RolloutMemory(1, bs, tags, S, I, E, dev)            # static entry both graphs work on
        self.cur.n = 1
        for e in (model.actor._engine, model.critic._engine):
            e.persistent_grads = True
        self._rollout_graph = self._update_graph = None
        self._warm, self._eager_updates = warmup, 0
        self.stats_sum = torch.zeros(10, device=dev)

    # -- rollout ------------------------------------------------------------------------------------------------
    def _rollout_eager(self):
        text, img, tgts = self.cur.text[0], self.cur.img[0], self.cur.tgts[0]
        entry = rollout(self.model, self.reward, text, img, tgts)
        for dst, src in zip((self.cur.state[0], self.cur.next_state[0], self.cur.scores[0], self.cur.rewards[0],
                             self.cur.value[0]), entry[:5]):
            dst.copy_(src.view(dst.shape))

    def rollout(self, text, img, tgts):
        """One rollout batch (fp32 tensors as the loader yields them, already on the device) into the memory."""
        self.cur.n = 0
        self.cur.stage(text, img, tgts)
        self.cur.n = 1
        if self._rollout_graph is None:
            self._rollout_graph = _capture(self._rollout_eager, self._warm)
        self._rollout_graph.replay()
        k = self.memory.n
        if k >= self.memory.capacity:
            raise RuntimeError("RolloutMemory is full: call update() first")
        for name in ("text", "img", "tgts", "state", "next_state", "scores", "rewards", "value"):
            getattr(self.memory, name)[k].copy_(getattr(self.cur, name)[0])
        self.memory.n += 1

    # -- update -------------------------------------------------------------------------------------------------
    def _update_eager(self):
        self.model.train()
        stats = update_batch(self.args, self.model, self.opt, self.copt, self.cur.entry(0), self.grad_sync)
        self.model.eval()
        self.stats_sum += stats

    def update(self, scheduler, critic_scheduler):
        """train_model over the stored batches (finetune/ppo.py:501-617); returns the ten averages and clears."""
        n = len(self.memory)
        self.stats_sum.zero_()
        for k in range(n):
            for name in ("text", "img", "tgts", "state", "next_state", "scores", "rewards", "value"):
                getattr(self.cur, name)[0].copy_(getattr(self.memory, name)[k])
            if self._update_graph is None and self._eager_updates < 2:
                # the first two updates of a run execute eagerly -- they ARE the updates of these entries, and they
                # allocate everything the captured graph needs (persistent .grad buffers, optimizer tables)
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    self._update_eager()
                torch.cuda.current_stream().wait_stream(side)
                self._eager_updates += 1
                continue
            if self._update_graph is None:
                torch.cuda.synchronize()
                self.opt.frozen_hyper = self.copt.frozen_hyper = True
                self._update_graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self._update_graph):       # records only: nothing executes during capture
                    self._update_eager()
            self.opt.update_hyper(); self.copt.update_hyper()
            self._update_graph.replay()
        total = self.stats_sum / max(n, 1)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            total = total / dist.get_world_size()
            dist.all_reduce(total)
        scheduler.step()
        critic_scheduler.step()
        self.memory.clear()
        return list(total.unbind(0))


def _capture(fn, warm):
    """Warm `fn` up on a side stream, then capture one call in a CUDA graph."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(warm):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        fn()
    return graph
