"""Parity at the BASELINE batch sizes (the tile plans, split-K choices and kernel-selection thresholds of the
benchmarked shapes), in TRAIN mode with replayed dropout masks, against goldens produced by the reference's own
train_model functions on full-size models (oracle/make_golden_r2.py):

  stage 3: ppo.sh:21            batch 24 x 2 tags  (actor / critic 48 items, reward 96 items)  -> stage3_bs24.pt
  stage 1: pointwise.sh:22,28   2 clips x 20 tags  (40 items)                                  -> stage12_full.pt
  stage 2: reward_pair_dataloader.sh:21  64 pairs, two forwards of 4-slot sequences (2 x 256 items)

bf16 compute: 2e-2 of each tensor's scale; rollout permutations bit-exact."""
import argparse
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from tests import golden_util, parity

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
TOL = 2e-2


def _margs():
    return argparse.Namespace(mode="reg", labels_num=3, seq_length=196, max_imgs=16, visual_feat_dim=768)


def _check_updates(test, named, before, gold, prefix="delta/"):
    """AdamW's first step moves every weight by ~lr * 3.16 * sign(g): compare the direction on the entries whose
    gradient is well above bf16 noise (sampled like the golden)."""
    rms = {n: gold["mnorm/" + n].item() / max(1.0, p.numel() ** 0.5) for n, p in named}
    top = max(rms.values())
    for n, p in named:
        ref = gold[prefix + n]
        if ref.abs().max().item() == 0 or rms[n] < 1e-3 * top or p.numel() < 64:
            continue          # mathematically-zero gradients (keys.bias, the actor's head.bias): the sign is noise
        got = golden_util.grad_sample(p.detach() - before[n], 4096).cpu()
        m = gold["m/" + n].reshape(-1)
        m = m if m.numel() == got.numel() else m[:got.numel()]
        big = m.abs() > 0.5 * m.abs().max()
        if big.sum() > 0:
            agree = (torch.sign(got[big]) == torch.sign(ref[big])).float().mean().item()
            parity.check(test, n + " [update sign disagreement on large-gradient entries]", 1.0 - agree, 0.02)


def test_stage3_step_batch24_train_mode_vs_reference_train_model():
    from lr2ppo_b200 import ppo
    gold = torch.load(os.path.join(GOLD, "stage3_bs24.pt"))
    c = golden_util.S3
    model = ppo.ActorCritic(_margs(), _margs())
    reward = ppo.Reward(_margs(), _margs())
    model.actor.load_state_dict(golden_util.make_state_dict("actor"), strict=True)
    model.critic.load_state_dict(golden_util.make_state_dict("critic"), strict=True)
    reward.load_state_dict(golden_util.make_state_dict("reward"), strict=True)
    model.cuda().eval(); reward.cuda().eval()
    hp = argparse.Namespace(is_master=False, mode="reg", kl_div_loss_weight=0.001, entropy_weight=0.001,
                            value_clip=0.5, learning_rate=c["lr"], critic_learning_rate=c["lr"], optimizer="adamw",
                            scheduler="linear", train_steps=c["train_steps"], warmup=c["warmup"],
                            fc1_grad_bf16=True)              # the configuration bench.py runs
    opt, copt, sch, csch = ppo.build_optimizer(hp, model)
    sch.step(); csch.step()                                  # as the generator: lr = 1e-3 / 10
    text, img, tgts = golden_util.stage3_inputs()
    text = text.cuda()
    img = img.cuda().unsqueeze(1).repeat(1, text.shape[1], 1, 1)          # finetune/ppo.py:831
    tgts = tgts.cuda()
    mem = ppo.rollout(model, reward, text, img, tgts)
    ro = gold["rollout"]
    test = "stage3 bs24x2 TRAIN (replayed masks)"
    assert torch.equal(mem[1].cpu(), ro["next_state"])                     # bit-exact permutations
    for nm, got, key in (("scores", mem[2], "action_scores"), ("rewards", mem[3], "rewards"), ("value", mem[4], "value")):
        parity.check(test, "rollout " + nm, parity.rel_err(got, ro[key]), TOL)
    # update on the REFERENCE's memory so both sides optimise the same objective
    mem_g = [mem[0], ro["next_state"].cuda(), ro["action_scores"].cuda(), ro["rewards"].cuda(), ro["value"].cuda(),
             text, img, tgts]
    model.actor._engine.dropout_seed = c["actor_seed"]
    model.critic._engine.dropout_seed = c["critic_seed"]
    before = {("actor." + n): p.detach().clone() for n, p in model.actor.named_parameters()}
    before.update({("critic." + n): p.detach().clone() for n, p in model.critic.named_parameters()})
    model.train()
    stats = ppo.train_model(hp, model, opt, copt, sch, csch, [mem_g], 0)
    names = ["policy_loss", "value_loss", "kl_penalty", "old_value", "value", "rewards_ori", "rewards", "advantages",
             "rank_loss", "entropy"]
    for nm, got, ref in zip(names, stats, gold["stats"]):
        # statistics that are differences of O(1) quantities (advantages = rewards - old_value, kl ~ 1e-6 ...) are
        # held to 2e-2 of the scale of their operands
        scale = max(abs(ref), 0.05)
        parity.check(test, "stat " + nm, abs(float(got) - ref) / scale, TOL)
    for tag, net, o in (("actor", model.actor, opt), ("critic", model.critic, copt)):
        named = [(f"{tag}.{n}", p) for n, p in net.named_parameters()]
        parity.check_param_tensors(f"{test} [{tag} exp_avg]", named, lambda p: o.state[p]["exp_avg"],
                                   lambda n: gold["m/" + n], lambda n: gold["mnorm/" + n].item(),
                                   lambda t: golden_util.grad_sample(t, 4096),
                                   elem_overrides=parity.CRITIC_TAIL_ELEM if tag == "critic" else None,
                                   norm_overrides=parity.CRITIC_TAIL_NORM if tag == "critic" else None)
        _check_updates(f"{test} [{tag}]", named, before, gold)


@pytest.mark.parametrize("stage", [1, 2])
def test_stage12_train_step_baseline_size_train_mode(stage):
    from lr2ppo_b200 import models, stages
    gold = torch.load(os.path.join(GOLD, "stage12_full.pt"))[f"stage{stage}"]
    c = golden_util.S12[stage]
    model = (models.Classifier if stage == 1 else models.PairClassifier)(_margs(), _margs())
    model.load_state_dict(golden_util.make_state_dict("actor" if stage == 1 else "reward"), strict=True)
    model = model.cuda().train()
    model._engine.dropout_seed = c["mask_seed"]
    # the shipped configurations: bf16 out_layer.fc1 gradient; stage 2 computes it with one GEMM over both passes
    hp = argparse.Namespace(learning_rate=golden_util.STEP_LR, optimizer="adamw", scheduler="constant",
                            fc1_grad_bf16=True, fc1_passes=stage)
    opt, sch = stages.build_optimizer(hp, model)
    named = list(model.named_parameters())
    before = {n: p.detach().clone() for n, p in named}
    text, img, tgts, chosen, reject = golden_util.stage12_inputs(stage)
    text = text.cuda()
    img = img.cuda().unsqueeze(1).repeat(1, text.shape[1], 1, 1)          # finetune/pointwise.py:544
    test = f"stage{stage} BASELINE size TRAIN (replayed masks)"
    if stage == 1:
        loss = stages.pointwise_train_model(hp, model, opt, sch, text, img, tgts.cuda())
    else:
        loss, acc = stages.reward_train_model(hp, model, opt, sch, text, img, tgts.cuda(), chosen.cuda(),
                                              reject.cuda())
        parity.check(test, "acc", abs(acc.item() - gold["acc"].item()), 2.0 / c["bs"])   # <= one borderline pair
        assert model._engine.dropout_seed == c["mask_seed"] + 2                           # two training forwards
    parity.check(test, "loss", abs(loss.item() - gold["loss"].item()) / abs(gold["loss"].item()), TOL)
    parity.check_param_tensors(test + " [exp_avg]", named, lambda p: opt.state[p]["exp_avg"],
                               lambda n: gold["m/" + n], lambda n: gold["mnorm/" + n].item(),
                               lambda t: golden_util.grad_sample(t, 4096),
                               elem_overrides=parity.pair_cancellation(8e-2) if stage == 2 else None)
    _check_updates(test, named, before, gold)
