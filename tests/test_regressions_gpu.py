"""Regression tests for defects found in review:

  * bf16 weight copies (engine.ShadowBank) must follow FusedAdamW updates on EVERY path, including optimizers built
    without ppo.attach_shadows (stage 1 / stage 2 / trad / standalone modules): the second forward after a step
    with lr > 0 must see the new weights;
  * GradSync.broadcast_params must not orphan the shadows registered with the optimizer;
  * the asynchronous checkpoint must hold the values of the moment save() was called, even if the tensors are
    updated in place immediately afterwards;
  * lr2_ppo_policy_loss takes [B, k] index lists with k != n (RankLoss, finetune/ppo.py:43-46) and never reads
    scores through an out-of-range index."""
import argparse
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from tests import golden_util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _margs():
    c = golden_util.FUSION_CFG
    return argparse.Namespace(mode="reg", labels_num=3, seq_length=c["seq_length"], max_imgs=c["max_imgs"],
                              visual_feat_dim=c["feat"])


def _plain_adamw(model, lr):
    """An optimizer built WITHOUT registering the engine's shadows (the round-1 stage-1/2 tests did exactly this)."""
    from lr2ppo_b200 import optim
    return optim.AdamW(optim.decay_groups(model.named_parameters()), lr=lr, correct_bias=False)


@pytest.mark.parametrize("stage", [1, 2])
def test_forward_sees_the_weights_after_an_unregistered_optimizer_step(stage):
    from lr2ppo_b200 import models, optim, stages
    cls = models.Classifier if stage == 1 else models.PairClassifier
    kind = "actor" if stage == 1 else "reward"
    model = cls(_margs(), _margs())
    model.load_state_dict(golden_util.make_state_dict(kind), strict=True)
    model = model.cuda().eval()
    text, img, tgts, chosen, reject = golden_util.stage_inputs(stage)
    text, img, tgts = text.cuda(), img.cuda(), tgts.cuda()

    def fwd(m):
        with torch.no_grad():
            return (m(text, img, None) if stage == 1 else m(text, img, tgts, chosen.cuda())).clone()

    y0 = fwd(model)
    opt = _plain_adamw(model, 1e-3)
    sch = optim.get_constant_schedule(opt)
    for _ in range(2):
        if stage == 1:
            stages.pointwise_train_model(None, model, opt, sch, text, img, tgts)
        else:
            stages.reward_train_model(None, model, opt, sch, text, img, tgts, chosen.cuda(), reject.cuda())
    y2 = fwd(model)
    assert (y2 - y0).abs().max().item() > 1e-2 * y0.abs().max().item()       # two lr = 1e-3 steps move the output
    fresh = cls(_margs(), _margs())
    fresh.load_state_dict({k: v.detach().cpu() for k, v in model.state_dict().items()}, strict=True)
    assert torch.equal(fwd(fresh.cuda().eval()), y2)                         # exactly what the fp32 masters give


def test_trad_forward_sees_the_weights_after_a_step():
    from lr2ppo_b200 import trad
    margs = argparse.Namespace(mode="reg", labels_num=5)
    model = trad.Actor(margs, margs)
    model.load_state_dict(golden_util.make_trad_state_dict("actor"), strict=True)
    model = model.cuda().eval()
    text, tgts, _ = golden_util.trad_inputs("actor")
    text, tgts = text.cuda(), tgts.cuda()
    with torch.no_grad():
        y0 = model(text, None, None).clone()
    opt = _plain_adamw(model, 1e-3)
    for _ in range(2):
        model.zero_grad()
        loss, _ = model(text, None, tgts)
        loss.backward()
        opt.step()
    with torch.no_grad():
        y2 = model(text, None, None).clone()
    assert (y2 - y0).abs().max().item() > 1e-2 * y0.abs().max().item()
    fresh = trad.Actor(margs, margs)
    fresh.load_state_dict({k: v.detach().cpu() for k, v in model.state_dict().items()}, strict=True)
    with torch.no_grad():
        assert torch.equal(fresh.cuda().eval()(text, None, None), y2)


def test_broadcast_params_keeps_the_registered_shadows(tmp_path):
    """world 1 process group: after build_optimizer -> broadcast_params (bench.py's order) the tensors the optimizer
    refreshes must still be the ones forward reads, and they must track two lr > 0 steps."""
    import torch.distributed as dist
    from lr2ppo_b200 import ppo
    from lr2ppo_b200.dist import GradSync
    if not dist.is_initialized():
        dist.init_process_group("gloo", init_method=f"file://{tmp_path}/pg", rank=0, world_size=1)
    model = ppo.ActorCritic(_margs(), _margs())
    model.actor.load_state_dict(golden_util.make_state_dict("actor"), strict=True)
    model.critic.load_state_dict(golden_util.make_state_dict("critic"), strict=True)
    model.cuda().eval()
    hp = argparse.Namespace(learning_rate=1e-3, critic_learning_rate=1e-3, optimizer="adamw", scheduler="constant",
                            fc1_grad_bf16=True)
    opt, copt, _, _ = ppo.build_optimizer(hp, model)
    eng = model.actor._engine
    w = model.actor.out_layer.fc1.weight
    registered = opt.shadow_of(w)
    bank_before = eng.bank
    with torch.no_grad():
        w.data.mul_(1.5)                                  # what a broadcast from rank 0 does: a `.data` write
    GradSync(1).broadcast_params(model)                   # world 1: values unchanged, but copies must be invalidated
    assert eng.bank is bank_before
    sh = eng.bank.get(w)
    assert sh.data_ptr() == registered.data_ptr()                       # same tensor the optimizer refreshes
    assert torch.equal(sh, w.detach().bfloat16())                        # and re-cast from the new values
    dist.destroy_process_group()


def test_async_checkpoint_is_not_torn_by_an_immediate_in_place_update(tmp_path):
    from lr2ppo_b200 import checkpoint
    n = 256 << 20                                                         # 1 GiB fp32: the D2H copy takes ~20+ ms
    t = torch.full((n,), 1.0, device="cuda")
    ck = checkpoint.AsyncCheckpointer()
    path = os.path.join(tmp_path, "state.bin")
    ck.save({"w": t}, path)
    t.add_(1.0)                                                           # the next optimizer step, enqueued at once
    t.mul_(3.0)
    ck.wait()
    got = torch.load(path)["w"]
    assert got.min().item() == 1.0 and got.max().item() == 1.0
    assert t[0].item() == 6.0


def test_policy_loss_index_lists_k_not_n_and_out_of_range():
    from lr2ppo_b200 import losses, ops
    for c in json.load(open(os.path.join(ROOT, "tests", "golden", "rows_r2.json")))["rank_loss_k"]:
        s = torch.tensor(c["scores"], device="cuda", requires_grad=True)
        idx = torch.tensor(c["order"], device="cuda")
        loss = losses.RankLoss(c["margin"])(s, idx)
        assert abs(loss.item() - c["loss"]) <= 1e-5 * max(1.0, abs(c["loss"])), (c["order"], loss.item(), c["loss"])
        loss.backward()
        ref = torch.tensor(c["dscores"])
        assert torch.allclose(s.grad.cpu(), ref, rtol=1e-4, atol=1e-7)
    s = torch.randn(8, 3, device="cuda")
    z = torch.zeros(8, device="cuda")
    bad = torch.tensor([[0, 3]] * 8, device="cuda")                       # 3 is outside [0, 3)
    out = ops.ppo_policy_loss(s, s, z, z, bad, 0.0, 0.0)
    torch.cuda.synchronize()
    assert torch.isnan(out["loss"]).item() and torch.isfinite(out["ds"]).all()
    with pytest.raises(Exception):
        ops.ppo_policy_loss(s, s, z[:5], z, bad, 0.0, 0.0)               # reward with the wrong number of rows


@pytest.mark.parametrize("kind", ["actor", "critic", "reward"])
def test_unrepeated_img_emb_is_broadcast_inside_the_gather(kind):
    """img_emb [bs, 1, I, E] (as the loader yields it) must give bit-identical results to the reference's
    img_emb.unsqueeze(1).repeat(1, tags, 1, 1) (finetune/ppo.py:831), forward and backward."""
    from lr2ppo_b200 import models
    cls = {"actor": models.Actor, "critic": models.Critic, "reward": models.Reward}[kind]
    model = cls(_margs(), _margs())
    model.load_state_dict(golden_util.make_state_dict(kind), strict=True)
    model = model.cuda().eval()
    text, img, tgts, index = golden_util.make_inputs(kind)
    text, one = text.cuda(), img[:, :1].contiguous().cuda()
    rep = one.repeat(1, text.shape[1], 1, 1)
    idx = None if index is None else index.cuda()

    def run(im):
        model.zero_grad(set_to_none=True)
        y = model.scores(text, im) if kind == "actor" else model(text, im, None, idx)
        y.sum().backward()
        return y.detach().clone(), {n: p.grad.clone() for n, p in model.named_parameters()}

    y1, g1 = run(one)
    y2, g2 = run(rep)
    assert torch.equal(y1, y2)
    for n in g1:
        assert torch.equal(g1[n], g2[n]), n
