"""CPU suite, round 2: the train-mode oracle restatement (oracle/fusion_ref.py with replayed masks) against the golden
produced by the reference's own modules with their nn.Dropout replaced by the same masks (fusion_train.pt), the Philox
restatement's fixed points, RankLoss on index lists shorter / longer than the score rows (rows_r2.json)."""
import json
import os

import numpy as np
import torch

from oracle import fusion_ref, philox, restate
from tests import golden_util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def test_rank_loss_with_k_different_from_n_vs_reference():
    for c in json.load(open(os.path.join(GOLD, "rows_r2.json")))["rank_loss_k"]:
        s = torch.tensor(c["scores"], requires_grad=True)
        loss, _ = restate.rank_loss(s, torch.tensor(c["order"]), c["margin"])
        assert abs(float(loss) - c["loss"]) <= 1e-7 * max(1.0, abs(c["loss"]))
        if loss.requires_grad:
            loss.backward()
            assert torch.allclose(s.grad, torch.tensor(c["dscores"]), rtol=1e-6, atol=1e-8)


def test_philox_mask_stream_properties_and_known_answers():
    m = philox.dropout_multiplier(12345, 2, 1 << 20, 0.1)
    assert philox.thresh16(0.1) == 6554 and abs(philox.scale16(0.1) - 65536.0 / (65536 - 6554)) < 1e-6
    keep = (m > 0).mean()
    assert abs(keep - (1 - 6554 / 65536)) < 2e-3                       # quantised keep probability
    assert set(np.unique(m).tolist()) == {0.0, float(np.float32(philox.scale16(0.1)))}
    assert abs(float(m.mean()) - 1.0) < 3e-3                            # E[multiplier] == 1
    # a pure function of (seed, site, index): prefixes agree, different seeds / sites do not
    assert np.array_equal(m[:4096], philox.dropout_multiplier(12345, 2, 4096, 0.1))
    assert not np.array_equal(m[:4096], philox.dropout_multiplier(12346, 2, 4096, 0.1))
    assert not np.array_equal(m[:4096], philox.dropout_multiplier(12345, 3, 4096, 0.1))
    # frozen fields (pin the restatement itself; tests/test_dropout_replay_gpu.py pins it to the CUDA kernel)
    f = philox.fields16(70001, 1, 2)
    assert f.shape == (2, 8) and f.dtype == np.uint16
    g = philox.fields16(70001, 1, 1, first_group=1)
    assert np.array_equal(f[1], g[0])
    big = philox.fields16((1 << 40) + 7, 6, 1, first_group=(1 << 33) + 5)     # 64-bit seed and counter halves used
    assert not np.array_equal(big, philox.fields16(7, 6, 1, first_group=5))


def test_train_mode_restatement_vs_reference_with_replayed_masks():
    """critic covers all six dropout sites (xit 1-3 on [items*196, .], xitt 4-6 on [bs*T, .])."""
    gold = torch.load(os.path.join(GOLD, "fusion_train.pt"))["critic"]
    sd = {k: v.requires_grad_(True) for k, v in golden_util.make_state_dict("critic").items()}
    text, img, tgts, index = golden_util.make_inputs("critic")
    bs, T = index.shape
    seed = golden_util.TRAIN_SEEDS["critic"]
    masks = philox.xit_masks(seed, 0, bs * T * 196, 768, 3072)
    for k, v in philox.xit_masks(seed, 3, bs * T, 768, 3072).items():
        masks[3 + k] = v
    logits = fusion_ref.critic_forward(sd, text, img, index, masks)
    assert torch.allclose(logits, gold["logits"], rtol=1e-5, atol=1e-5)
    eval_logits = fusion_ref.critic_forward(sd, text, img, index)
    assert (eval_logits - logits).abs().max() > 1e-3                    # the masks do change the function
    (logits * golden_util.out_grad("critic", logits.numel())).sum().backward()
    for name, p in sd.items():
        ref = gold["grad/" + name]
        got = p.grad if ref.numel() == p.grad.numel() else golden_util.grad_sample(p.grad)
        scale = max(ref.abs().max().item(), gold["gnorm/" + name].item() / p.numel() ** 0.5)
        assert (got.reshape(-1) - ref.reshape(-1)).abs().max().item() <= 2e-5 * scale + 1e-9, name
