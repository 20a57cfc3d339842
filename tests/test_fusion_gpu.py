"""Fusion model (Actor / Critic / Reward) on the CUDA engine vs the reference-generated golden outputs
(tests/golden/fusion.pt: reference modules run on CPU fp32 with seed-generated weights).
bf16 compute: tolerance 2e-2 relative to the tensor's scale (north_star)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from tests import golden_util, parity

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = torch.load(os.path.join(ROOT, "tests", "golden", "fusion.pt"))
TOL = 2e-2


def _args():
    import argparse
    c = golden_util.FUSION_CFG
    return argparse.Namespace(mode="reg", labels_num=3, seq_length=c["seq_length"], max_imgs=c["max_imgs"],
                              visual_feat_dim=c["feat"])


def _build(kind):
    from lr2ppo_b200 import models
    cls = {"actor": models.Actor, "critic": models.Critic, "reward": models.Reward}[kind]
    model = cls(_args(), _args())
    model.load_state_dict(golden_util.make_state_dict(kind), strict=True)   # checkpoint-key contract
    return model.cuda().eval()


def _rel(d, ref):
    d, ref = d.detach().float().cpu(), ref.float()
    return ((d - ref).abs().max() / ref.abs().max().clamp_min(1e-12)).item()


@pytest.mark.parametrize("kind", ["actor", "critic", "reward"])
def test_forward_backward_vs_reference_golden(kind):
    model = _build(kind)
    text, img, tgts, index = golden_util.make_inputs(kind)
    text, img, tgts = text.cuda(), img.cuda(), tgts.cuda()
    if kind == "actor":
        loss, logits = model(text, img, tgts)
    else:
        logits = model(text, img, tgts, index.cuda())
    gold = GOLD[kind]
    assert logits.dtype == torch.float32 and logits.shape == gold["logits"].shape
    parity.check(f"fusion[{kind}] bs2x2 eval", "logits", _rel(logits, gold["logits"]), TOL)
    gw = golden_util.out_grad(kind, logits.numel()).cuda()
    (logits * gw).sum().backward()
    named = list(model.named_parameters())
    for name, p in named:
        assert p.grad is not None and p.grad.dtype == torch.float32, name
    parity.check_param_tensors(f"fusion[{kind}] bs2x2 eval", named, lambda p: p.grad, lambda n: gold["grad/" + n],
                               lambda n: gold["gnorm/" + n].item(), golden_util.grad_sample)


def test_actor_no_grad_and_tgts_none():
    model = _build("actor")
    text, img, tgts, _ = golden_util.make_inputs("actor")
    with torch.no_grad():
        logits = model(text.cuda(), img.cuda(), None)
    assert _rel(logits, GOLD["actor"]["logits"]) < TOL
    assert all(p.grad is None for p in model.parameters())


def test_train_mode_dropout_runs_and_is_seed_deterministic():
    model = _build("critic").train()
    text, img, tgts, index = golden_util.make_inputs("critic")
    text, img, tgts, index = text.cuda(), img.cuda(), tgts.cuda(), index.cuda()
    eng = model._engine
    a, _ = eng.forward(text, img, index, train=True, save=False, seed=5)
    b, _ = eng.forward(text, img, index, train=True, save=False, seed=5)
    c, _ = eng.forward(text, img, index, train=True, save=False, seed=6)
    assert torch.equal(a, b) and not torch.equal(a, c)
    out = model(text, img, tgts, index)
    out.sum().backward()
    for n, p in model.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), n


def test_gradient_accumulates_over_two_forwards():
    # stage 2 runs the model twice (chosen / reject) before one backward (reward_pair_dataloader.py:351-358)
    model = _build("reward")
    text, img, tgts, index = golden_util.make_inputs("reward")
    text, img, tgts, index = text.cuda(), img.cuda(), tgts.cuda(), index.cuda()
    s1 = model(text, img, tgts, index)
    s1.sum().backward()
    g1 = {n: p.grad.clone() for n, p in model.named_parameters()}
    model.zero_grad(set_to_none=True)
    c = model(text, img, tgts, index)
    r = model(text, img, tgts, index)
    (c.sum() + r.sum()).backward()
    for n, p in model.named_parameters():
        ref = 2 * g1[n]
        err = (p.grad - ref).abs().max().item() / max(ref.abs().max().item(), 1e-20)
        assert err < 1e-2, (n, err)
