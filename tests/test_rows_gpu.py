"""GPU parity of the memory-bound row kernels against the oracle and the reference-generated golden vectors.
Integer / index / NDCG results: bit-exact.  Floating point: 1e-5 (fp32), tolerance written per test."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from lr2ppo_b200 import ops
from oracle import restate

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
ROWS = json.load(open(os.path.join(GOLD, "rows.json")))
UPD = json.load(open(os.path.join(GOLD, "ppo_update.json")))
KS = [1, 3, 5, 10, 20, 100000000]


def cu(x, dtype=torch.float32):
    return torch.tensor(x, dtype=dtype, device="cuda")


def test_ndcg_golden_bit_exact():
    for c in ROWS["ndcg"]:
        out, order = ops.ndcg_at_k(cu([c["scores"]]), cu([c["labels"]], torch.int64), c["ks"], want_order=True)
        assert order[0].tolist() == c["order"]
        assert out[0].cpu().numpy().tobytes() == np.asarray(c["ndcg"], dtype=np.float32).tobytes()


@pytest.mark.parametrize("N", [1, 2, 16, 20, 33, 64, 128, 256, 512, 1024, 1500])
@pytest.mark.parametrize("nlab", [3, 5])
def test_ndcg_vs_oracle_bit_exact(N, nlab):
    rng = np.random.default_rng(N * 10 + nlab)
    B = 64
    scores = rng.standard_normal((B, N)).astype(np.float32)     # tie-free with probability ~1
    labels = rng.integers(0, nlab, (B, N))
    ref, ref_order = restate.ndcg_at_k(scores, labels, KS, want_order=True)
    out, order = ops.ndcg_at_k(cu(scores), cu(labels, torch.int64), KS, want_order=True)
    assert np.array_equal(order.cpu().numpy(), ref_order)
    assert out.cpu().numpy().tobytes() == ref.tobytes()


def test_ndcg_ties_ragged_and_all_zero():
    rng = np.random.default_rng(5)
    B, N = 32, 40
    scores = rng.integers(0, 4, (B, N)).astype(np.float32)       # heavy ties -> stable order (lower index first)
    labels = rng.integers(0, 3, (B, N))
    lens = rng.integers(1, N + 1, B).astype(np.int32)
    ref, ref_order = restate.ndcg_at_k(scores, labels, KS, lens=lens, want_order=True)
    out, order = ops.ndcg_at_k(cu(scores), cu(labels, torch.int64), KS, lens=cu(lens, torch.int32), want_order=True)
    assert np.array_equal(order.cpu().numpy(), ref_order)
    assert out.cpu().numpy().tobytes() == ref.tobytes()
    out = ops.ndcg_at_k(cu(scores), torch.zeros((B, N), dtype=torch.int64, device="cuda"), KS)
    assert (out == 1).all()


@pytest.mark.parametrize("N", [7, 32, 33, 64, 200, 1024])
@pytest.mark.parametrize("kind", ["short_runs", "long_runs", "signed_zero_and_ties", "mixed_lens"])
def test_ndcg_truncated_key_repair_is_exact(N, kind):
    """The 32-bit warp kernel sorts on the top 22 bits of the score key and repairs the order of colliding keys
    (odd-even rounds, then a full-key shared-memory sort when runs are long).  Scores that differ only in their LOW
    bits exercise every branch; the result must still be the stable descending order of the exact fp32 scores."""
    rng = np.random.default_rng(N * 7 + len(kind))
    B = 96
    if kind == "short_runs":          # pairs / triples inside one 2^-13 relative bucket, plus ordinary scores
        base = rng.standard_normal((B, N)).astype(np.float32)
        twin = rng.integers(0, N, (B, N))
        near = np.take_along_axis(base, twin, 1) * (1 + rng.integers(-3, 4, (B, N)).astype(np.float32) * np.float32(2 ** -21))
        scores = np.where(rng.random((B, N)) < 0.3, near, base).astype(np.float32)
    elif kind == "long_runs":         # hundreds of scores inside one bucket: forces the bail-out path
        scores = (1.0 + rng.integers(0, 60, (B, N)) * np.float32(2 ** -20)).astype(np.float32)
        scores[::2] *= -1
    elif kind == "signed_zero_and_ties":
        scores = rng.integers(-2, 3, (B, N)).astype(np.float32)
        scores[scores == 0] = np.where(rng.random((scores == 0).sum()) < 0.5, -0.0, 0.0)
    else:
        scores = (rng.standard_normal((B, N)) * np.float32(1e-3) + 5).astype(np.float32)   # one exponent, dense
    labels = rng.integers(0, 5, (B, N))
    lens = rng.integers(1, N + 1, B).astype(np.int32) if kind == "mixed_lens" else None
    ref, ref_order = restate.ndcg_at_k(scores, labels, KS, lens=lens, want_order=True)
    out, order = ops.ndcg_at_k(cu(scores), cu(labels, torch.int64), KS,
                               lens=None if lens is None else cu(lens, torch.int32), want_order=True)
    assert np.array_equal(order.cpu().numpy(), ref_order)
    assert out.cpu().numpy().tobytes() == ref.tobytes()


@pytest.mark.parametrize("N", [48, 300, 1024])
def test_ndcg_general_labels_ragged_unsorted_cuts(N):
    # labels outside [0, 62] in every other query (second register sort of the int64 label keys) next to
    # histogram-path queries in the same launch; ragged lengths; cuts unsorted, duplicated, zero and oversized
    rng = np.random.default_rng(N)
    B = 16
    scores = rng.standard_normal((B, N)).astype(np.float32)
    scores[:, : N // 3] = np.round(scores[:, : N // 3])          # ties
    labels = rng.integers(0, 63, (B, N))
    for q in range(0, B, 2):
        labels[q, rng.integers(0, N, 4)] = [63, -2, 70, 2 ** 40]
    lens = rng.integers(1, N + 1, B).astype(np.int32)
    lens[0] = N
    ks = [7, 1, 100000000, 7, 0, 3]
    ref, ref_order = restate.ndcg_at_k(scores, labels, ks, lens=lens, want_order=True)
    out, order = ops.ndcg_at_k(cu(scores), cu(labels, torch.int64), ks, lens=cu(lens, torch.int32), want_order=True)
    assert np.array_equal(order.cpu().numpy(), ref_order)
    assert out.cpu().numpy().tobytes() == ref.tobytes()


def test_ndcg_full_size_properties():
    # config-5 extreme (B=4096, N=1024): size-independent properties
    g = torch.Generator(device="cuda").manual_seed(0)
    B, N = 4096, 1024
    scores = torch.randn(B, N, generator=g, device="cuda")
    labels = torch.randint(0, 3, (B, N), generator=g, device="cuda")
    out, order = ops.ndcg_at_k(scores, labels, KS, want_order=True)
    assert (order.sort(dim=1).values == torch.arange(N, device="cuda")).all()         # a permutation per query
    assert (torch.gather(scores, 1, order).diff(dim=1) <= 0).all()                    # sortedness
    assert ((out >= 0) & (out <= 1.0000001)).all()
    perfect = ops.ndcg_at_k(labels.float(), labels, KS)                                # ranking by the labels -> 1
    assert (perfect == 1).all()
    sub = restate.ndcg_at_k(scores[:8].cpu().numpy(), labels[:8].cpu().numpy(), KS)
    assert out[:8].cpu().numpy().tobytes() == sub.tobytes()


def test_rollout_golden_and_sampler_bit_exact():
    for c in ROWS["rollout"]:
        ns = ops.ppo_rollout(cu(c["scores"]), None, 2)
        assert ns.tolist() == c["next_state"]
        perm, _ = ops.rank_sample(cu(c["scores"]), greedy=True)
        assert perm.tolist() == [r[2:] for r in c["next_state"]]
    rng = np.random.default_rng(3)
    for B, n in [(24, 2), (256, 20), (1000, 80)]:
        s = rng.standard_normal((B, n)).astype(np.float32)
        u = rng.random((B, n)).astype(np.float32)
        ref_perm, ref_lp = restate.rank_sample(s, u)
        perm, lp = ops.rank_sample(cu(s), cu(u))
        assert np.array_equal(perm.cpu().numpy(), ref_perm)                            # bit-exact permutations
        assert lp.cpu().numpy().tobytes() == ref_lp.tobytes()                          # deterministic exp/log
        st = np.stack([rng.permutation(n) for _ in range(B)])
        ref_ns, ref_order = restate.ppo_rollout(s, st, 2)
        ns, order = ops.ppo_rollout(cu(s), cu(st, torch.int64), 2, want_order=True)
        assert np.array_equal(ns.cpu().numpy(), ref_ns) and np.array_equal(order.cpu().numpy(), ref_order)


def test_gae_scan():
    rng = np.random.default_rng(4)
    for B, T in [(24, 1), (7, 5), (33, 32), (16, 100), (3, 1000)]:
        r = rng.standard_normal((B, T)).astype(np.float32)
        v = rng.standard_normal((B, T + 1)).astype(np.float32)
        nd = (rng.random((B, T)) > 0.1).astype(np.float32)
        ref_adv, ref_ret = restate.gae(r, v, 0.99, 0.95, nd)
        adv, ret = ops.gae_scan(cu(r), cu(v), 0.99, 0.95, cu(nd))
        np.testing.assert_allclose(adv.cpu().numpy(), ref_adv, rtol=1e-5, atol=1e-5)    # fp32, scan re-association
        np.testing.assert_allclose(ret.cpu().numpy(), ref_ret, rtol=1e-5, atol=1e-5)
    r = rng.standard_normal((24, 1)).astype(np.float32)
    v = np.concatenate([rng.standard_normal((24, 1)).astype(np.float32), np.zeros((24, 1), np.float32)], 1)
    adv, _ = ops.gae_scan(cu(r), cu(v), 0.99, 0.95)
    assert np.array_equal(adv.cpu().numpy()[:, 0], r[:, 0] - v[:, 0])                  # T=1 -> r - V exactly


def test_ppo_update_vs_reference_train_model_golden():
    for c in UPD:
        pi = cu(c["next_state"], torch.int64)[:, -2:].contiguous()
        r = ops.ppo_policy_loss(cu(c["s_new"]), cu(c["s_old"]), cu(c["reward"]), cu(c["v_old"]), pi, c["w_kl"],
                                c["w_ent"])
        st = c["stats"]
        tol = 1e-5
        assert abs(r["loss"].item() - st["policy_loss"]) <= tol * max(1, abs(st["policy_loss"]))
        assert abs(r["rank_loss"].item() - st["rank_loss"]) <= tol
        assert abs(r["kl"].mean().item() - st["kl"]) <= tol
        assert abs(r["entropy"].mean().item() - st["entropy"]) <= tol
        assert abs(r["adv"].mean().item() - st["advantages"]) <= tol
        assert abs(r["reward_adj"].mean().item() - st["rewards"]) <= tol
        np.testing.assert_allclose(r["ds"].cpu().numpy(), np.asarray(c["ds"], np.float32), rtol=1e-4, atol=1e-7)
        vl, dv = ops.clipped_value_loss(cu(c["v_new"]), r["reward_adj"], cu(c["v_old"]), c["value_clip"])
        assert abs(vl.item() - st["value_loss"]) <= tol
        np.testing.assert_allclose(dv.cpu().numpy(), np.asarray(c["dv"], np.float32), rtol=1e-5, atol=1e-8)


def test_policy_loss_general_n_vs_oracle():
    g = torch.Generator().manual_seed(11)
    for B, n in [(24, 2), (100, 5), (300, 20)]:
        s = torch.randn(B, n, generator=g) * 0.3
        s_old = s + torch.randn(B, n, generator=g) * 0.1
        reward = torch.randn(B, generator=g); v_old = torch.randn(B, generator=g)
        pi = torch.stack([torch.randperm(n, generator=g) for _ in range(B)])
        sr = s.clone().requires_grad_(True)
        ref = restate.ppo_policy_loss(sr, s_old, reward, v_old, pi, 0.05, 0.02)
        ref["loss"].backward()
        out = ops.ppo_policy_loss(s.cuda(), s_old.cuda(), reward.cuda(), v_old.cuda(), pi.cuda(), 0.05, 0.02)
        assert abs(out["loss"].item() - ref["loss"].item()) <= 1e-5 * max(1, abs(ref["loss"].item()))
        assert out["hinge_cnt"].item() == ref["hinge_cnt"].item()
        np.testing.assert_allclose(out["ds"].cpu().numpy(), sr.grad.numpy(), rtol=1e-4, atol=1e-7)


def test_losses_golden():
    for c in ROWS["value_loss"]:
        loss, dv = ops.clipped_value_loss(cu(c["v"]), cu(c["ret"]), cu(c["v_old"]), c["clip"])
        assert abs(loss.item() - c["loss"]) <= 1e-5 * max(1, abs(c["loss"]))
        np.testing.assert_allclose(dv.cpu().numpy(), np.asarray(c["dv"], np.float32), rtol=1e-5, atol=1e-8)
    for c in ROWS["pair_hinge"]:
        loss, acc, dc, dr = ops.pair_hinge_loss(cu(c["chosen"]), cu(c["reject"]), c["margin"])
        assert abs(loss.item() - c["loss"]) <= 1e-5 and abs(acc.item() - c["acc"]) <= 1e-6
        np.testing.assert_allclose(dc.cpu().numpy(), np.asarray(c["dchosen"], np.float32), rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(dr.cpu().numpy(), np.asarray(c["dreject"], np.float32), rtol=1e-6, atol=1e-9)
    for c in ROWS["smooth_l1"]:
        loss, dl = ops.smooth_l1_loss(cu(c["logits"]), cu(c["tgt"], torch.int64), c["beta"])
        assert abs(loss.item() - c["loss"]) <= 1e-5
        np.testing.assert_allclose(dl.cpu().numpy(), np.asarray(c["dlogits"], np.float32), rtol=1e-5, atol=1e-8)
    for c in ROWS["rank_loss"]:
        s = cu(c["scores"])
        B, n = s.shape
        # RankLoss alone: zero KL/entropy weights, advantage forced >= eps so the given order is used as is
        out = ops.ppo_policy_loss(s, s, torch.ones(B, device="cuda"), torch.zeros(B, device="cuda"),
                                  cu(c["order"], torch.int64), 0.0, 0.0, margin=c["margin"])
        assert abs(out["rank_loss"].item() - c["loss"]) <= 1e-6


def test_adamw_golden_and_large():
    from lr2ppo_b200.optim import FusedAdamW
    for c in ROWS["adamw"]:
        p = torch.nn.Parameter(cu(c["p0"]))
        opt = FusedAdamW([{"params": [p], "weight_decay": c["wd"]}], lr=c["lr"], correct_bias=False)
        for g in c["grads"]:
            p.grad = cu(g)
            opt.step()
        np.testing.assert_allclose(p.detach().cpu().numpy(), np.asarray(c["p"], np.float32), rtol=1e-5, atol=1e-8)
        st = opt.state_for(p)
        np.testing.assert_allclose(st["exp_avg"].cpu().numpy(), np.asarray(c["m"], np.float32), rtol=1e-5, atol=1e-9)
        np.testing.assert_allclose(st["exp_avg_sq"].cpu().numpy(), np.asarray(c["v"], np.float32), rtol=1e-5,
                                   atol=1e-12)
    # multi-tensor, odd sizes, bf16 grads + bf16 shadow weights, vs the oracle restatement
    g = torch.Generator().manual_seed(2)
    shapes = [(3072, 768), (768,), (5, 7), (1,), (4099,)]
    ps = [torch.randn(s, generator=g) * 0.02 for s in shapes]
    gs = [torch.randn(s, generator=g) * 0.01 for s in shapes]
    params = [torch.nn.Parameter(p.clone().cuda()) for p in ps]
    opt = FusedAdamW([{"params": params[:1] + params[2:3], "weight_decay": 0.01},
                      {"params": params[1:2] + params[3:], "weight_decay": 0.0}], lr=1e-3, correct_bias=False,
                     shadow_bf16=True)
    ref_p = [p.clone() for p in ps]; ref_m = [torch.zeros_like(p) for p in ps]; ref_v = [torch.zeros_like(p) for p in ps]
    wds = [0.01, 0.0, 0.01, 0.0, 0.0]
    for step in range(3):
        for p, gr in zip(params, gs):
            p.grad = (gr * (step + 1)).cuda()
        opt.step()
        for i in range(len(ps)):
            restate.adamw_step(ref_p[i], gs[i] * (step + 1), ref_m[i], ref_v[i], 1e-3, wds[i])
    for p, rp in zip(params, ref_p):
        np.testing.assert_allclose(p.detach().cpu().numpy(), rp.numpy(), rtol=1e-5, atol=1e-8)
        sh = opt.shadow_of(p)
        assert torch.equal(sh, p.detach().to(torch.bfloat16))


def test_adamw_two_phase_step_is_bit_identical():
    """step(first=..., between=...) (data-parallel overlap: the big tensor is updated while the all-reduce of the
    rest is in flight) must give exactly the bits of a single-launch step, and must call `between` exactly once
    before any other tensor is touched."""
    from lr2ppo_b200.optim import FusedAdamW
    g = torch.Generator().manual_seed(5)
    shapes = [(768,), (64, 4099), (3072, 768), (5, 7), (9000,)]
    ps = [torch.randn(s, generator=g) * 0.02 for s in shapes]
    gs = [torch.randn(s, generator=g) * 0.01 for s in shapes]

    def make():
        params = [torch.nn.Parameter(p.clone().cuda()) for p in ps]
        opt = FusedAdamW([{"params": [params[1], params[2], params[3]], "weight_decay": 0.01},
                          {"params": [params[0], params[4]], "weight_decay": 0.0}], lr=1e-3, correct_bias=False,
                         shadow_bf16=True)
        return params, opt

    pa, oa = make()
    pb, ob = make()
    calls = []
    for step in range(3):
        for p, q, gr in zip(pa, pb, gs):
            p.grad = (gr * (step + 1)).cuda()
            q.grad = (gr * (step + 1)).cuda()
        oa.step()
        snap = [q.detach().clone() for q in pb]

        def between():
            torch.cuda.synchronize()
            calls.append(step)
            # only the `first` tensor may have changed so far
            for i, (q, s0) in enumerate(zip(pb, snap)):
                assert torch.equal(q.detach(), s0) == (i != 2), i

        ob.step(first={id(pb[2])}, between=between)
    assert calls == [0, 1, 2]
    for p, q in zip(pa, pb):
        assert torch.equal(p.detach(), q.detach())
        assert torch.equal(oa.state_for(p)["exp_avg_sq"], ob.state_for(q)["exp_avg_sq"])
        assert torch.equal(oa.shadow_of(p), ob.shadow_of(q))


def test_adamw_row_windows_partition_the_update():
    """Row-sharded optimizer (dist.GradSync shard_fc1): two 'ranks' that each own half of the big tensor's chunks
    together produce exactly the single-optimizer update; chunks of the other rank are left untouched."""
    from lr2ppo_b200.optim import FusedAdamW
    g = torch.Generator().manual_seed(8)
    shapes = [(768,), (64, 8192), (5, 7)]          # the big one: 128 chunks of 4096
    ps = [torch.randn(s, generator=g) * 0.02 for s in shapes]
    gs = [torch.randn(s, generator=g) * 0.01 for s in shapes]

    def make(window=None):
        params = [torch.nn.Parameter(p.clone().cuda()) for p in ps]
        opt = FusedAdamW([{"params": [params[1], params[2]], "weight_decay": 0.01},
                          {"params": [params[0]], "weight_decay": 0.0}], lr=1e-3, correct_bias=False,
                         shadow_bf16=True)
        if window is not None:
            opt.set_window(params[1], window, 2)
        return params, opt

    full, of = make()
    r0, o0 = make(0)
    r1, o1 = make(1)
    for step in range(2):
        for params in (full, r0, r1):
            for p, gr in zip(params, gs):
                p.grad = (gr * (step + 1)).cuda()
        of.step()
        o0.step(first={id(r0[1])}, between=lambda: None)
        o1.step()
    half = 32
    assert torch.equal(r0[1].detach()[:half], full[1].detach()[:half])
    assert torch.equal(r1[1].detach()[half:], full[1].detach()[half:])
    assert torch.equal(r0[1].detach()[half:], ps[1].cuda()[half:])          # not owned: untouched
    assert torch.equal(r1[1].detach()[:half], ps[1].cuda()[:half])
    assert torch.equal(o0.shadow_of(r0[1])[:half], of.shadow_of(full[1])[:half])
    for i in (0, 2):
        assert torch.equal(r0[i].detach(), full[i].detach()) and torch.equal(r1[i].detach(), full[i].detach())
