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


def test_rollout_memory_ring_semantics():
    """ppo.RolloutMemory host logic (bf16 sources: no kernel involved): slots fill in order, entries come back in
    insertion order with the reference's 8-field layout (finetune/ppo.py:878-883), the image set is stored
    un-repeated, a full ring refuses another batch and clear() rewinds it."""
    import pytest
    from lr2ppo_b200 import ppo
    bs, tags, S, I, E = 3, 2, 4, 5, 8
    mem = ppo.RolloutMemory(2, bs, tags, S, I, E, "cpu")
    g = torch.Generator().manual_seed(0)
    batches = []
    for k in range(2):
        text = torch.randn(bs, tags, S, E, generator=g).bfloat16()
        img = torch.randn(bs, I, E, generator=g).bfloat16()                       # as the loader yields it: [bs, I, E]
        tgts = torch.randint(0, 3, (bs, tags), generator=g)
        t, i, y = mem.stage(text, img, tgts)
        assert i.shape == (bs, 1, I, E) and torch.equal(t, text) and torch.equal(i[:, 0], img) and torch.equal(y, tgts)
        state = torch.arange(tags).repeat(bs, 1)
        next_state = torch.randint(0, tags, (bs, 2 + tags), generator=g)
        scores, rewards, value = torch.randn(bs, tags, generator=g), torch.randn(bs, 1, generator=g), torch.randn(bs, generator=g)
        mem.commit([state, next_state, scores, rewards, value, t, i, y])
        batches.append((state, next_state, scores, rewards.view(-1), value, text, img.unsqueeze(1), tgts))
    assert len(mem) == 2
    with pytest.raises(RuntimeError):
        mem.stage(*[b for b in (batches[0][5], batches[0][6], batches[0][7])])
    for got, want in zip(mem, batches):
        assert len(got) == 8
        for a, b in zip(got, want):
            assert torch.equal(a, b)
    assert mem.bytes_per_entry() == bs * ((tags * S + I) * E * 2 + tags * 8 * 2 + (2 + tags) * 8 + tags * 4 + 8)
    mem.clear()
    assert len(mem) == 0 and len(list(mem)) == 0
    mem.stage(batches[1][5], batches[1][6], batches[1][7])                          # [bs, 1, I, E] accepted as well
    assert torch.equal(mem.text[0], batches[1][5])


def test_branch_streams_are_cuda_only_and_env_gated(monkeypatch):
    """The multi-branch step (ppo._branch_stream / _reward_stream) never engages off CUDA, and LR2_DUAL_STREAM=0 /
    LR2_REWARD_STREAM=0 switch it off: the single-stream order of the reference remains one variable away."""
    from lr2ppo_b200 import ppo
    cpu = torch.device("cpu")
    assert ppo._branch_stream(cpu) is None and ppo._reward_stream(cpu) is None
    monkeypatch.setenv("LR2_DUAL_STREAM", "0")
    assert ppo._branch_stream(torch.device("cuda", 0)) is None and ppo._reward_stream(torch.device("cuda", 0)) is None
