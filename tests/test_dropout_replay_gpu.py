"""Train-mode (dropout live) parity.  The reference's update runs with three nn.Dropout(0.1) per XiT block live
(finetune/xit.py:26-41, model.train() at finetune/ppo.py:886).  The CUDA path's masks are a pure function of
(seed, site, element index): (1) lr2_dropout_bf16 and the GEMM / LayerNorm-backward epilogues draw exactly the stream
oracle/philox.py restates (bit for bit); (2) with FusionEngine.dropout_seed fixed, a TRAIN-mode forward + backward of
Actor / Critic / Reward matches the golden produced by the reference's own modules whose nn.Dropout were replaced by a
multiply with those masks (tests/golden/fusion_train.pt, oracle/make_golden_r2.py).  bf16 compute: 2e-2 of scale."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import philox
from tests import golden_util, parity

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 2e-2


@pytest.mark.parametrize("seed,site,rows,cols,p", [(70001, 1, 392, 768, 0.1), (70002, 5, 4, 3072, 0.1),
                                                   ((1 << 40) + 99, 2, 96, 3072, 0.25), (5, 3, 7, 8, 0.5)])
def test_dropout_kernel_draws_the_restated_philox_stream(seed, site, rows, cols, p):
    from lr2ppo_b200 import ops
    ones = torch.ones(rows, cols, dtype=torch.bfloat16, device="cuda")
    got = ops.dropout(ones, p, seed, site).float().cpu().numpy().reshape(-1)
    ref = philox.dropout_multiplier(seed, site, rows * cols, p)
    ref_bf16 = torch.from_numpy(ref).to(torch.bfloat16).float().numpy()        # the kernel stores bf16
    assert np.array_equal(got, ref_bf16)


def test_gemm_epilogue_and_layernorm_backward_use_the_same_stream():
    """out = dropout(A W^T + b) + res with res = 0 and b = 0 must be (A W^T) * mask elementwise."""
    from lr2ppo_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    a = torch.randn(256, 768, generator=g, device="cuda").bfloat16()
    w = (torch.randn(768, 768, generator=g, device="cuda") * 0.05).bfloat16()
    zero_b = torch.zeros(768, device="cuda")
    zero_r = torch.zeros(256, 768, dtype=torch.bfloat16, device="cuda")
    seed, site, p = 424242, 3, 0.1
    plain = ops.gemm(a, w, epilogue=ops.EPI_BIAS, bias=zero_b, out_dtype=torch.float32)
    dropped = ops.gemm(a, w, epilogue=ops.EPI_BIAS_DROP_RES, bias=zero_b, aux=zero_r, drop_p=p, seed=seed, site=site)
    mask = torch.from_numpy(philox.dropout_multiplier(seed, site, 256 * 768, p)).view(256, 768).cuda()
    assert torch.equal((dropped == 0), (mask == 0) | (plain.bfloat16() == 0))
    ref = (plain * mask).bfloat16().float()
    assert (dropped.float() - ref).abs().max().item() <= 2 ** -7 * ref.abs().max().item()
    # LayerNorm backward emits the masked copy of dx with the same stream
    x = torch.randn(256, 768, generator=g, device="cuda").bfloat16()
    dy = torch.randn(256, 768, generator=g, device="cuda").bfloat16()
    gamma = torch.ones(768, device="cuda"); beta = torch.zeros(768, device="cuda")
    _, st = ops.layernorm_fwd(x, gamma, beta, 1e-5)
    dx, dxm, _, _ = ops.layernorm_bwd(dy, x, gamma, st, 1e-5, drop_p=p, seed=seed, site=site, want_masked=True)
    assert torch.equal(dxm == 0, (mask == 0) | (dx == 0))


def _build(kind):
    import argparse
    from lr2ppo_b200 import models
    c = golden_util.FUSION_CFG
    a = argparse.Namespace(mode="reg", labels_num=3, seq_length=c["seq_length"], max_imgs=c["max_imgs"],
                           visual_feat_dim=c["feat"])
    model = {"actor": models.Actor, "critic": models.Critic, "reward": models.Reward}[kind](a, a)
    model.load_state_dict(golden_util.make_state_dict(kind), strict=True)
    return model.cuda()


@pytest.mark.parametrize("kind", ["actor", "critic", "reward"])
def test_train_mode_forward_backward_vs_reference_with_replayed_masks(kind):
    gold = torch.load(os.path.join(ROOT, "tests", "golden", "fusion_train.pt"))[kind]
    model = _build(kind).train()
    model._engine.dropout_seed = golden_util.TRAIN_SEEDS[kind]
    text, img, tgts, index = golden_util.make_inputs(kind)
    if kind == "actor":
        _, logits = model(text.cuda(), img.cuda(), tgts.cuda())
    else:
        logits = model(text.cuda(), img.cuda(), tgts.cuda(), index.cuda())
    assert model._engine.dropout_seed == golden_util.TRAIN_SEEDS[kind] + 1       # one training forward consumed
    test = f"fusion[{kind}] bs2x2 TRAIN (replayed masks)"
    parity.check(test, "logits", parity.rel_err(logits, gold["logits"]), TOL)
    # the eval-mode golden differs from the train-mode one by far more than the tolerance: the masks matter
    ev = torch.load(os.path.join(ROOT, "tests", "golden", "fusion.pt"))[kind]["logits"]
    assert parity.rel_err(ev, gold["logits"]) > 5 * TOL
    (logits * golden_util.out_grad(kind, logits.numel()).cuda()).sum().backward()
    parity.check_param_tensors(test, list(model.named_parameters()), lambda p: p.grad, lambda n: gold["grad/" + n],
                               lambda n: gold["gnorm/" + n].item(), golden_util.grad_sample)
