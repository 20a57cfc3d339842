"""CPU suite: the oracle restatements reproduce the golden vectors generated from the imported reference
(oracle/make_golden.py), and the C-ABI library loads and exports every symbol include/lr2ppo_b200.h declares."""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch

from oracle import restate

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
ROWS = json.load(open(os.path.join(GOLD, "rows.json")))
UPD = json.load(open(os.path.join(GOLD, "ppo_update.json")))


def f32(x):
    return np.asarray(x, dtype=np.float32)


def test_ndcg_oracle_bit_exact_vs_reference():
    for c in ROWS["ndcg"]:
        scores, labels = f32(c["scores"])[None], np.asarray(c["labels"], dtype=np.int64)[None]
        out, order = restate.ndcg_at_k(scores, labels, c["ks"], want_order=True)
        assert order[0].tolist() == c["order"]                      # bit-exact indices
        assert out[0].tobytes() == f32(c["ndcg"]).tobytes()          # bit-exact fp32 NDCG
        pred = labels[0][order[0]]
        ideal = np.sort(labels[0])[::-1]
        assert restate.ndcg_python(pred, ideal, c["ks"]).tobytes() == f32(c["ndcg"]).tobytes()


def test_ndcg_known_answers():
    # SURVEY.md §8c known-answer vectors
    ks = [1, 3, 5, 10, 20, 100000000]
    out = restate.ndcg_at_k(f32([[6, 5, 4, 3, 2, 1]]), np.array([[2, 0, 1, 0, 2, 1]]), ks)[0]
    np.testing.assert_allclose(out, [1.0, 0.649014771, 0.800306618, 0.861474216, 0.861474216, 0.861474216], rtol=1e-7)
    out = restate.ndcg_at_k(f32([[3, 2, 1]]), np.zeros((1, 3), dtype=np.int64), ks)[0]
    assert out.tolist() == [1.0] * 6
    # ragged: lens cuts the list
    a = restate.ndcg_at_k(f32([[3, 1, 2, 9, 9]]), np.array([[0, 2, 1, 2, 2]]), ks, lens=[3])[0]
    b = restate.ndcg_at_k(f32([[3, 1, 2]]), np.array([[0, 2, 1]]), ks)[0]
    assert a.tobytes() == b.tobytes()


def test_rank_loss_and_known_answers():
    for c in ROWS["rank_loss"]:
        loss, _ = restate.rank_loss(torch.tensor(c["scores"]), torch.tensor(c["order"]), c["margin"])
        assert abs(float(loss) - c["loss"]) <= 1e-7 * max(1.0, abs(c["loss"]))
    loss, _ = restate.rank_loss(torch.tensor([[.30, .10], [.20, .25], [.50, .50]]),
                                torch.tensor([[0, 1], [0, 1], [1, 0]]), 0.01)
    assert abs(float(loss) - 0.034999996423721313) < 1e-9
    v = restate.clipped_value_loss(torch.tensor([.9, -.2]), torch.tensor([.5, .1]), torch.tensor([.1, 0.]), .5)
    assert abs(float(v) - 0.1249999925494194) < 1e-9


def test_value_hinge_smoothl1_vs_reference():
    for c in ROWS["value_loss"]:
        v = torch.tensor(c["v"], requires_grad=True)
        loss = restate.clipped_value_loss(v, torch.tensor(c["ret"]), torch.tensor(c["v_old"]), c["clip"])
        loss.backward()
        assert abs(float(loss) - c["loss"]) < 1e-6
        np.testing.assert_allclose(v.grad.numpy(), f32(c["dv"]), rtol=1e-5, atol=1e-7)
    for c in ROWS["pair_hinge"]:
        loss, acc = restate.pair_hinge_loss(torch.tensor(c["chosen"]), torch.tensor(c["reject"]), c["margin"])
        assert abs(float(loss) - c["loss"]) < 1e-6 and abs(float(acc) - c["acc"]) < 1e-7
    for c in ROWS["smooth_l1"]:
        loss = restate.smooth_l1(torch.tensor(c["logits"]), torch.tensor(c["tgt"]), c["beta"])
        assert abs(float(loss) - c["loss"]) < 1e-6


def test_rollout_vs_reference_and_greedy_sampler_reduces_to_sort():
    for c in ROWS["rollout"]:
        s = f32(c["scores"])
        ns, order = restate.ppo_rollout(s, None, 2)
        assert ns.tolist() == c["next_state"]
        perm, _ = restate.rank_sample(s, greedy=True)               # greedy sampler == torch.sort(desc)
        assert perm.tolist() == order.tolist()


def test_sampler_properties():
    rng = np.random.default_rng(0)
    s = rng.standard_normal((64, 9)).astype(np.float32)
    u = rng.random((64, 9)).astype(np.float32)
    perm, lp = restate.rank_sample(s, u)
    assert (np.sort(perm, axis=1) == np.arange(9)).all()            # always a permutation
    assert np.isfinite(lp).all() and (lp <= 0).all()
    # u -> 0 always takes the first remaining label in index order
    perm0, _ = restate.rank_sample(s, np.zeros_like(u))
    assert (perm0 == np.arange(9)).all()
    # Plackett-Luce log-probability matches a float64 evaluation
    b = 3
    ref = 0.0
    rem = list(range(9))
    for t in range(9):
        z = s[b, rem].astype(np.float64)
        ref += z[rem.index(perm[b, t])] - np.log(np.exp(z - z.max()).sum()) - z.max()
        rem.remove(perm[b, t])
    assert abs(lp[b] - ref) < 1e-4


def test_gae_reductions():
    rng = np.random.default_rng(1)
    r = rng.standard_normal((5, 1)).astype(np.float32)
    v = np.concatenate([rng.standard_normal((5, 1)).astype(np.float32), np.zeros((5, 1), np.float32)], 1)
    adv, ret = restate.gae(r, v, 0.99, 0.95)
    assert (adv[:, 0] == r[:, 0] - v[:, 0]).all()                   # T=1, V_next=0 -> r - V (finetune/ppo.py:560)
    T = 37
    r = rng.standard_normal((4, T)).astype(np.float32)
    v = rng.standard_normal((4, T + 1)).astype(np.float32)
    adv, ret = restate.gae(r, v, 0.97, 0.9)
    ref = np.zeros((4, T))
    nxt = np.zeros(4)
    for t in reversed(range(T)):
        nxt = r[:, t] + 0.97 * v[:, t + 1] - v[:, t] + 0.97 * 0.9 * nxt
        ref[:, t] = nxt
    np.testing.assert_allclose(adv, ref, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(ret, ref + v[:, :T], rtol=1e-5, atol=1e-5)


def test_ppo_update_restatement_vs_reference_train_model():
    """restate.ppo_policy_loss / clipped_value_loss == finetune/ppo.py:train_model run on stub networks."""
    for c in UPD:
        s = torch.tensor(c["s_new"], requires_grad=True)
        pi = torch.tensor(c["next_state"])[:, -2:]
        r = restate.ppo_policy_loss(s, torch.tensor(c["s_old"]), torch.tensor(c["reward"]), torch.tensor(c["v_old"]),
                                    pi, c["w_kl"], c["w_ent"])
        r["loss"].backward()
        st = c["stats"]
        assert abs(float(r["loss"]) - st["policy_loss"]) < 1e-6
        assert abs(float(r["rank_loss"]) - st["rank_loss"]) < 1e-6
        assert abs(float(r["kl"].mean()) - st["kl"]) < 1e-6
        assert abs(float(r["entropy"].mean()) - st["entropy"]) < 1e-6
        assert abs(float(r["adv"].mean()) - st["advantages"]) < 1e-6
        np.testing.assert_allclose(s.grad.numpy(), f32(c["ds"]), rtol=1e-5, atol=1e-8)
        v = torch.tensor(c["v_new"], requires_grad=True)
        vl = restate.clipped_value_loss(v, r["reward_adj"].detach(), torch.tensor(c["v_old"]), c["value_clip"])
        vl.backward()
        assert abs(float(vl) - st["value_loss"]) < 1e-6
        np.testing.assert_allclose(v.grad.numpy(), f32(c["dv"]), rtol=1e-5, atol=1e-8)


def test_adamw_and_schedule_vs_reference():
    for c in ROWS["adamw"]:
        p = torch.tensor(c["p0"]); m = torch.zeros_like(p); v = torch.zeros_like(p)
        for g in c["grads"]:
            restate.adamw_step(p, torch.tensor(g), m, v, c["lr"], c["wd"])
        assert p.numpy().tobytes() == f32(c["p"]).tobytes()          # same op sequence -> bit-exact on CPU
        assert m.numpy().tobytes() == f32(c["m"]).tobytes()
        assert v.numpy().tobytes() == f32(c["v"]).tobytes()
    sc = ROWS["schedule"]
    lrs = [sc["base_lr"] * restate.linear_schedule_lambda(i, sc["warmup"], sc["total"]) for i in range(6)]
    np.testing.assert_allclose(lrs, sc["lrs"], rtol=1e-12)
    assert lrs[0] == 0.0                                             # lr is exactly 0 before the first scheduler.step()


def test_tencent_layernorm_vs_reference():
    c = ROWS["tencent_ln"]
    x = torch.tensor(c["x"], requires_grad=True)
    gamma = torch.tensor(c["gamma"], requires_grad=True); beta = torch.tensor(c["beta"], requires_grad=True)
    y = restate.tencent_layernorm(x, gamma, beta, 1e-6)
    (y * torch.tensor(c["gy"])).sum().backward()
    np.testing.assert_allclose(y.detach().numpy(), f32(c["y"]), rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(x.grad.numpy(), f32(c["dx"]), rtol=1e-4, atol=1e-5)


def test_c_abi_library_exports_every_declared_symbol():
    from lr2ppo_b200 import _lib
    header = open(os.path.join(ROOT, "include", "lr2ppo_b200.h")).read()
    declared = set(re.findall(r"\b(lr2_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert os.path.exists(_lib.LIB_PATH), "build the library first (__graft_entry__.build())"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.lr2_abi_version() == 1
    lib.lr2_last_error_string.restype = ctypes.c_char_p
    assert lib.lr2_last_error_string(-4).startswith(b"device is not")


@pytest.mark.parametrize("kind", ["actor", "critic", "reward"])
def test_fusion_restatement_vs_reference_modules(kind):
    """oracle/fusion_ref.py == reference Actor/Critic/Reward (finetune/ppo.py:196-350) on seeded tensors."""
    from oracle import fusion_ref
    from tests import golden_util
    gold = torch.load(os.path.join(GOLD, "fusion.pt"))
    sd = golden_util.make_state_dict(kind)
    text, img, tgts, index = golden_util.make_inputs(kind)
    with torch.no_grad():
        if kind == "actor":
            logits = fusion_ref.actor_forward(sd, text, img)
        else:
            logits = fusion_ref.critic_forward(sd, text, img, index)
    ref = gold[kind]["logits"]
    assert (logits - ref).abs().max() <= 1e-5 * ref.abs().max().clamp_min(1.0), (logits, ref)


@pytest.mark.parametrize("kind", ["actor", "critic", "reward"])
def test_trad_restatement_vs_reference_modules(kind):
    """oracle/fusion_ref.trad_* == reference ppo_trad Actor/Critic/Reward (finetune/ppo_trad.py:142-283)."""
    from oracle import fusion_ref
    from tests import golden_util
    gold = torch.load(os.path.join(GOLD, "trad.pt"))
    sd = golden_util.make_trad_state_dict(kind)
    text, tgts, index = golden_util.trad_inputs(kind)
    with torch.no_grad():
        logits = fusion_ref.trad_actor_forward(sd, text) if kind == "actor" else \
            fusion_ref.trad_critic_forward(sd, text, index)
    ref = gold[kind]["logits"]
    assert (logits - ref).abs().max() <= 1e-5 * ref.abs().max().clamp_min(1.0)


@pytest.mark.parametrize("kind", ["vit", "roberta"])
def test_tower_restatement_vs_reference_build_model(kind):
    """oracle/tower_ref.py == reference build_model(ViT-B/16 | RoBERTa-base).embedding/encoder on seeded tensors."""
    from oracle import tower_ref
    from tests import golden_util
    gold = torch.load(os.path.join(GOLD, "tower.pt"))[kind]
    sd = golden_util.make_tower_state_dict(gold["names"], golden_util.TOWER_SEEDS[kind])
    src, seg = golden_util.tower_inputs(kind)
    with torch.no_grad():
        emb = tower_ref.embed_patch(sd, src) if kind == "vit" else tower_ref.embed_word(sd, src, seg)
        hid = tower_ref.encoder(sd, emb, seg, 12, 12, pre_ln=(kind == "vit"))
    ref = gold["hidden"]
    assert (hid - ref).abs().max() <= 2e-5 * ref.abs().max()
