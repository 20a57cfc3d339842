"""ppo.evaluate: packing the (clip, tag) items of several clips into one actor forward gives the
same scores and NDCG as the reference's one-clip-per-forward loop (finetune/ppo.py:620-681)."""
import argparse
import types

import pytest
import torch

pytestmark = pytest.mark.gpu

from lr2ppo_b200 import ppo


def test_batched_evaluate_matches_per_clip_evaluate():
    torch.manual_seed(3)
    margs = argparse.Namespace(mode="reg", labels_num=3, seq_length=196, max_imgs=16, visual_feat_dim=768)
    with torch.device("cuda"):
        actor = ppo.Actor(margs, margs)
    with torch.no_grad():
        for n, p in actor.named_parameters():
            if "gamma" not in n and "beta" not in n:
                p.normal_(0, 0.02)
    model = types.SimpleNamespace(actor=actor, eval=lambda: actor.eval())
    g = torch.Generator().manual_seed(9)
    loader = []
    for tags in (3, 7, 20, 1, 12, 5, 9):
        loader.append((torch.randn(1, tags, 196, 768, generator=g), torch.randn(1, 16, 768, generator=g),
                       torch.randint(0, 3, (1, tags), generator=g)))

    def run(cap):
        args = argparse.Namespace(model=model, device=torch.device("cuda"), is_master=True, eval_items=cap)
        captured = {}
        orig = ppo.AverageNDCGMeter.batch_ndcg

        def spy(self, scores, labels, lens=None, want_order=False):
            captured["scores"] = scores.clone()
            return orig(self, scores, labels, lens=lens, want_order=want_order)
        ppo.AverageNDCGMeter.batch_ndcg = spy
        try:
            ndcg = ppo.evaluate(args, loader, 0)
        finally:
            ppo.AverageNDCGMeter.batch_ndcg = orig
        return float(ndcg), captured["scores"]

    n1, s1 = run(1)            # one clip per forward (reference behaviour)
    n2, s2 = run(24)           # a few clips per forward
    n3, s3 = run(1000)         # everything in one forward
    fin = torch.isfinite(s1)
    assert torch.equal(fin, torch.isfinite(s2)) and torch.equal(fin, torch.isfinite(s3))
    scale = s1[fin].abs().max().item()
    assert (s1[fin] - s2[fin]).abs().max().item() < 2e-2 * scale
    assert (s1[fin] - s3[fin]).abs().max().item() < 2e-2 * scale
    # packing changes GEMM tile plans, not the function: same ranking => bit-identical NDCG (parity against the
    # reference's own evaluate output is tests/test_dropin_gpu.py::test_evaluate_vs_the_reference_evaluate_functions)
    o1, o2, o3 = (torch.argsort(s, dim=1, descending=True, stable=True) for s in (s1, s2, s3))
    if torch.equal(o1, o2):
        assert n1 == n2
    if torch.equal(o1, o3):
        assert n1 == n3
    assert abs(n1 - n2) < 0.05 and abs(n1 - n3) < 0.05
