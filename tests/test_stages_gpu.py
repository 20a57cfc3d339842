"""Stage-1 / stage-2 training steps and the MSLR ("trad") models on the CUDA path vs goldens produced by the
reference's own `train_model` functions and modules (oracle/make_golden.py: trad.pt, stage12.pt).
bf16 compute: 2e-2 of each tensor's scale."""
import argparse
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from tests import golden_util, parity

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
TOL = 2e-2


def _rel(d, ref):
    d, ref = d.detach().float().cpu(), ref.float()
    return ((d - ref).abs().max() / ref.abs().max().clamp_min(1e-12)).item()


def _check_grads(test, named, getter, gold, prefix, nprefix, **kw):
    parity.check_param_tensors(test, named, getter, lambda n: gold[prefix + n], lambda n: gold[nprefix + n].item(),
                               lambda t: golden_util.grad_sample(t, 4096), **kw)


@pytest.mark.parametrize("kind", ["actor", "critic", "reward"])
def test_trad_models_vs_reference_golden(kind):
    from lr2ppo_b200 import trad
    gold = torch.load(os.path.join(GOLD, "trad.pt"))[kind]
    args = argparse.Namespace(mode="reg", labels_num=5)
    model = {"actor": trad.Actor, "critic": trad.Critic, "reward": trad.Reward}[kind](args, args)
    model.load_state_dict(golden_util.make_trad_state_dict(kind), strict=True)
    model = model.cuda().eval()
    text, tgts, index = golden_util.trad_inputs(kind)
    if kind == "actor":
        _, logits = model(text.cuda(), None, tgts.cuda())
    else:
        logits = model(text.cuda(), None, tgts.cuda(), index.cuda())
    parity.check(f"trad[{kind}]", "logits", _rel(logits, gold["logits"]), TOL)
    (logits * golden_util.out_grad(kind, logits.numel()).cuda()).sum().backward()
    # reward: the 4-slot index [0, 1, pi0, pi1] feeds each item twice to xitt, whose query / key gradients are then
    # differences of nearly equal contributions of duplicate rows (measured 0.065)
    _check_grads(f"trad[{kind}]", list(model.named_parameters()), lambda p: p.grad, gold, "grad/", "gnorm/",
                 elem_overrides={"xitt.": 8e-2} if kind == "reward" else None)


@pytest.mark.parametrize("stage", [1, 2])
def test_stage_train_step_vs_reference_train_model(stage):
    from lr2ppo_b200 import models, optim, stages
    gold = torch.load(os.path.join(GOLD, "stage12.pt"))[f"stage{stage}"]
    cfg = golden_util.FUSION_CFG
    args = argparse.Namespace(mode="reg", labels_num=3, seq_length=cfg["seq_length"], max_imgs=cfg["max_imgs"],
                              visual_feat_dim=cfg["feat"])
    model = (models.Classifier if stage == 1 else models.PairClassifier)(args, args)
    model.load_state_dict(golden_util.make_state_dict("actor" if stage == 1 else "reward"), strict=True)
    model = model.cuda().train()
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    named = list(model.named_parameters())
    no_decay = ["bias", "gamma", "beta"]
    groups = [{"params": [p for n, p in named if not any(nd in n for nd in no_decay)], "weight_decay": 0.01},
              {"params": [p for n, p in named if any(nd in n for nd in no_decay)], "weight_decay": 0.0}]
    opt = optim.AdamW(groups, lr=golden_util.STEP_LR, correct_bias=False)
    sch = optim.get_constant_schedule(opt)
    text, img, tgts, chosen, reject = golden_util.stage_inputs(stage)
    before = {n: p.detach().clone() for n, p in named}
    if stage == 1:
        loss = stages.pointwise_train_model(args, model, opt, sch, text.cuda(), img.cuda(), tgts.cuda())
    else:
        loss, acc = stages.reward_train_model(args, model, opt, sch, text.cuda(), img.cuda(), tgts.cuda(),
                                              chosen.cuda(), reject.cuda())
        assert abs(acc.item() - gold["acc"].item()) < 1e-6
    parity.check(f"stage{stage} train step", "loss", abs(loss.item() - gold["loss"].item()) / abs(gold["loss"].item()), TOL)
    _check_grads(f"stage{stage} train step [exp_avg]", named, lambda p: opt.state[p]["exp_avg"], gold, "m/", "mnorm/",
                 elem_overrides=parity.pair_cancellation(0.12) if stage == 2 else None)   # 3 pairs only
    # parameter update direction: delta = -lr * m/(sqrt(v)+eps) - lr*wd*p ; compare on the sampled entries
    rms = {n: gold["mnorm/" + n].item() / max(1.0, p.numel() ** 0.5) for n, p in named}
    top = max(rms.values())
    for n, p in named:
        if rms[n] < 1e-4 * top:
            continue                                                  # mathematically-zero gradients (keys.bias)
        ref = gold["delta/" + n]
        got = golden_util.grad_sample(p.detach() - before[n], 4096).cpu()
        big = ref.abs() > 0.5 * ref.abs().max()                      # entries whose gradient is well above bf16 noise
        if big.sum() > 0 and gold["mnorm/" + n].item() > 0:
            assert (torch.sign(got[big]) == torch.sign(ref[big])).float().mean() > 0.98, n


@pytest.mark.parametrize("kind", ["video", "proj"])
def test_api_modules_vs_reference(kind):
    """VideoTransformer / ProjectionLayer (imported-but-unused reference modules): same constructor, same
    state_dict keys, outputs and gradients vs the reference's own modules (tests/golden/api.pt)."""
    from lr2ppo_b200.project_embedding import ProjectionLayer
    from lr2ppo_b200.video_transformer import VideoTransformer
    gold = torch.load(os.path.join(GOLD, "api.pt"))[kind]
    model = VideoTransformer(**golden_util.VIDEO_CFG) if kind == "video" else ProjectionLayer(768, 512, 0.2)
    model.load_state_dict(golden_util.make_api_state_dict(gold["names"], golden_util.API_SEEDS[kind]), strict=True)
    model = model.cuda().eval()
    x = golden_util.api_input(kind).cuda().requires_grad_(True)
    y = model(x)
    assert y.shape == gold["y"].shape
    parity.check(f"api[{kind}]", "y", _rel(y, gold["y"]), TOL)
    (y * golden_util.out_grad("critic", y.numel()).view_as(y).cuda()).sum().backward()
    parity.check(f"api[{kind}]", "dx", _rel(x.grad, gold["dx"]), TOL)
    _check_grads(f"api[{kind}]", list(model.named_parameters()), lambda p: p.grad, gold, "grad/", "gnorm/")
