"""features.FeatureExtractor: towers -> clean_feat.h5-shaped tensors on the fly."""
import argparse

import pytest
import torch

pytestmark = pytest.mark.gpu

from lr2ppo_b200 import tower
from lr2ppo_b200.features import FeatureExtractor
from tests.test_tower_gpu import ROBERTA, VIT

VOCAB = 1000


def _tower(cfg, layers=2):
    args = argparse.Namespace(**dict(cfg, layers_num=layers))
    m = tower.build_model(args, vocab_size=VOCAB)
    g = torch.Generator().manual_seed(layers)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "gamma" not in n and "beta" not in n:
                p.copy_(torch.randn(p.shape, generator=g) * 0.05)
    return m.cuda().eval()


def test_feature_shapes_padding_and_cyclic_image_pad():
    fx = FeatureExtractor(_tower(VIT), _tower(ROBERTA))
    g = torch.Generator().manual_seed(1)
    tokens = torch.randint(5, VOCAB, (5, 64), generator=g).cuda()
    frames = torch.randn(3, 3, 224, 224, generator=g).cuda()
    text_emb, img_emb = fx(frames, tokens)
    assert text_emb.shape == (5, 196, 768) and text_emb.dtype == torch.float32
    assert img_emb.shape == (1, 3, 768) and img_emb.dtype == torch.float32
    # the 64 real positions do not see the 132 padded keys (seg = 0): same as running the tower on 64 tokens
    direct = fx.text(tokens, None, torch.ones(5, 64, dtype=torch.int64, device="cuda"))
    err = (text_emb[:, :64] - direct).abs().max().item() / direct.abs().max().item()
    assert err < 2e-2, err
    # ragged lengths: positions beyond a tag's length are masked for every query
    lens = torch.tensor([64, 10, 33, 1, 50])
    t2 = fx.text_features(tokens, lens)
    d2 = fx.text(tokens[1:2, :10], None, torch.ones(1, 10, dtype=torch.int64, device="cuda"))
    assert (t2[1, :10] - d2[0]).abs().max().item() / d2.abs().max().item() < 2e-2
    # image CLS vectors == position 0 of the ViT output
    hid = fx.vit(frames, None, torch.ones(3, 197, dtype=torch.int64, device="cuda"))
    assert torch.equal(img_emb[0], hid[:, 0, :])
    # cyclic pad (finetune/pointwise.py:149-154) without shuffle, and truncation
    padded = fx.pad_images(img_emb, shuffle=False)
    assert padded.shape == (16, 768)
    for i in range(16):
        assert torch.equal(padded[i], img_emb[0, i % 3])
    many = torch.randn(1, 20, 768, device="cuda")
    assert torch.equal(fx.pad_images(many, shuffle=False), many[0, :16])
    shuffled = fx.pad_images(img_emb, shuffle=True)
    assert sorted(shuffled[:3].sum(1).tolist()) == pytest.approx(sorted(img_emb[0].sum(1).tolist()))


def test_features_vs_the_reference_towers_golden():
    """Full 12-layer ViT-B/16 and RoBERTa-base towers with the weights and inputs of tests/golden/tower.pt (hidden
    states produced by the reference's own tencentpretrain build_model): the feature extractor's outputs must be those
    hidden states -- CLS position per keyframe for img_emb; for text_emb the real-token positions of every tag, which
    padding the sequence to the fusion model's 196 positions (masked keys) must not change.  bf16: 2e-2 of scale."""
    from tests import golden_util, parity
    from tests.test_tower_gpu import GOLD, _build
    vit, gold_v = _build("vit")
    rob, gold_r = _build("roberta")
    fx = FeatureExtractor(vit, rob)
    frames, _ = golden_util.tower_inputs("vit")
    img_emb = fx.image_features(frames.cuda())
    assert img_emb.shape == (1, frames.shape[0], 768)
    parity.check("features vs reference towers", "img_emb (ViT CLS)", parity.rel_err(img_emb[0], gold_v["hidden"][:, 0, :]), 2e-2)
    tokens, seg = golden_util.tower_inputs("roberta")
    lens = seg.sum(dim=1)
    text_emb = fx.text_features(tokens.cuda(), lens.cuda())
    assert text_emb.shape == (tokens.shape[0], 196, 768)
    scale = gold_r["hidden"].abs().max().item()
    worst = 0.0
    for i, n in enumerate(lens.tolist()):
        d = (text_emb[i, :n].float().cpu() - gold_r["hidden"][i, :n]).abs().max().item() / scale
        worst = max(worst, d)
    parity.check("features vs reference towers", "text_emb (RoBERTa, real-token positions)", worst, 2e-2)
