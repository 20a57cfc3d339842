"""TencentPretrain towers (ViT-B/16, RoBERTa-base via build_model) on the CUDA path vs the golden produced by the
reference's own build_model (oracle/make_golden.py tower).  bf16 compute through 12 layers: 2e-2 of tensor scale
on the hidden states, 5e-2 on sampled gradient entries / 2e-2 on gradient norms."""
import argparse
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from tests import golden_util, parity

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = torch.load(os.path.join(ROOT, "tests", "golden", "tower.pt"))

VIT = dict(emb_size=768, feedforward_size=3072, hidden_size=768, hidden_act="gelu", heads_num=12, layers_num=12,
           dropout=0.1, max_seq_length=197, embedding=["patch", "pos"], remove_embedding_layernorm=True,
           encoder="transformer", mask="fully_visible", layernorm_positioning="pre", image_height=224,
           image_width=224, patch_size=16)                       # models/vit/base-16-224_config.json
ROBERTA = dict(emb_size=768, feedforward_size=3072, hidden_size=768, hidden_act="gelu", heads_num=12, layers_num=12,
               max_seq_length=514, dropout=0.1, embedding=["word", "pos", "seg"], encoder="transformer",
               mask="fully_visible")                              # models/xlm-roberta/base_config.json


def _build(kind):
    from lr2ppo_b200 import tower
    args = argparse.Namespace(**(VIT if kind == "vit" else ROBERTA))
    model = tower.build_model(args, vocab_size=golden_util.TOWER_VOCAB)
    gold = GOLD[kind]
    sd = golden_util.make_tower_state_dict(gold["names"], golden_util.TOWER_SEEDS[kind])
    model.load_state_dict(sd, strict=True)                        # reference key names (embedding.*, encoder.*)
    return model.cuda().eval(), gold


@pytest.mark.parametrize("kind", ["vit", "roberta"])
def test_tower_forward_backward_vs_reference_golden(kind):
    model, gold = _build(kind)
    src, seg = golden_util.tower_inputs(kind)
    hidden = model(src.cuda(), None, seg.cuda()).float()
    ref = gold["hidden"]
    err = ((hidden.detach().cpu() - ref).abs().max() / ref.abs().max()).item()
    parity.check(f"tower[{kind}] 12 layers", "hidden", err, 2e-2)
    gw = golden_util.out_grad("actor", hidden.numel()).view_as(hidden).cuda()
    (hidden * gw).sum().backward()
    named = [(n, p) for n, p in model.named_parameters() if ("gnorm/" + n) in gold]
    assert len(named) == len(gold["names"])
    # 12 layers of bf16 activations: the attention key / query projections of the last layers reach 0.044
    parity.check_param_tensors(f"tower[{kind}] 12 layers", named, lambda p: p.grad, lambda n: gold["grad/" + n],
                               lambda n: gold["gnorm/" + n].item(), golden_util.grad_sample, elem_tol=5e-2)


def test_tower_train_mode_runs():
    model, _ = _build("vit")
    model.train()
    src, seg = golden_util.tower_inputs("vit")
    out = model(src.cuda(), None, seg.cuda()).float()
    out.sum().backward()
    for n, p in model.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), n
