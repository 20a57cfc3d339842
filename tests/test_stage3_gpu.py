"""Stage-3 LR2PPO step (rollout + update) on the CUDA engine vs the CPU oracle restatement
(oracle/stage3_ref.py, pinned to the reference by tests/golden/*).  The reference update runs with dropout 0.1
live; parity is defined with dropout off (model.eval()), as SURVEY.md §7 prescribes.
bf16 compute: 2e-2 relative to each tensor's scale; permutations / indices bit-exact."""
import argparse

import pytest
import torch

pytestmark = pytest.mark.gpu

from tests import golden_util, parity
from oracle import stage3_ref

TOL = 2e-2


def _margs():
    c = golden_util.FUSION_CFG
    return argparse.Namespace(mode="reg", labels_num=3, seq_length=c["seq_length"], max_imgs=c["max_imgs"],
                              visual_feat_dim=c["feat"])


def _hp(fused, bf16_grad=False):
    return argparse.Namespace(fc1_grad_bf16=bf16_grad, learning_rate=1e-5, critic_learning_rate=2e-5, optimizer="adamw", scheduler="constant",
                              train_steps=1000, warmup=0.1, kl_div_loss_weight=0.001, entropy_weight=0.001,
                              value_clip=0.5, mode="reg", fused_fc1=fused)


def _rel(d, ref):
    d, ref = d.detach().float().cpu(), ref.detach().float().cpu()
    return ((d - ref).abs().max() / ref.abs().max().clamp_min(1e-12)).item()


def _build_gpu(sds):
    from lr2ppo_b200 import ppo
    model = ppo.ActorCritic(_margs(), _margs())
    reward = ppo.Reward(_margs(), _margs())
    model.actor.load_state_dict(sds["actor"]); model.critic.load_state_dict(sds["critic"])
    reward.load_state_dict(sds["reward"])
    return model.cuda().eval(), reward.cuda().eval()


@pytest.fixture(scope="module")
def sds():
    return {k: golden_util.make_state_dict(k) for k in ("actor", "critic", "reward")}


def test_stage3_step_vs_cpu_oracle(sds):
    from lr2ppo_b200 import ppo
    g = torch.Generator().manual_seed(5)
    bs = 6
    text = torch.randn(bs, 2, 196, 768, generator=g)
    img = torch.randn(bs, 1, 16, 768, generator=g).repeat(1, 2, 1, 1)
    tgts = torch.randint(0, 3, (bs, 2), generator=g)
    # ---- CPU oracle
    ra = stage3_ref.RefModel(sds["actor"]); rc = stage3_ref.RefModel(sds["critic"])
    rr = stage3_ref.RefModel(sds["reward"], trainable=False)
    mem_ref, out_ref = stage3_ref.step(ra, rc, rr, text, img, 1e-5, 2e-5)
    # ---- CUDA engine
    model, reward = _build_gpu(sds)
    hp = _hp(False, bf16_grad=True)              # the configuration bench.py runs: bf16 out_layer.fc1 gradient
    opt, copt, sch, csch = ppo.build_optimizer(hp, model)
    mem = ppo.rollout(model, reward, text.cuda(), img.cuda(), tgts.cuda())
    state, next_state, scores, rewards, value = mem[:5]
    assert torch.equal(next_state.cpu(), mem_ref[1])                      # bit-exact permutations
    for nm, got, ref in (("scores", scores, mem_ref[2]), ("rewards", rewards, mem_ref[3]), ("value", value, mem_ref[4])):
        parity.check("stage3 step bs6", nm, _rel(got, ref), TOL)
    # update on the ORACLE's memory so both sides optimise the same objective
    mem_g = [mem_ref[0].cuda(), mem_ref[1].cuda(), mem_ref[2].cuda(), mem_ref[3].cuda(), mem_ref[4].cuda(),
             text.cuda(), img.cuda(), tgts.cuda()]
    stats = ppo.update_batch(hp, model, opt, copt, mem_g)
    names = ["policy_loss", "value_loss"]
    for i, n in enumerate(names):
        parity.check("stage3 step bs6", n, abs(stats[i].item() - out_ref[n].item()) / max(1e-3, abs(out_ref[n].item())), TOL)
    # first Adam moment = 0.1 * gradient: linear in the gradient, compared on every parameter
    for tag, net, ref, o in (("actor", model.actor, ra, opt), ("critic", model.critic, rc, copt)):
        parity.check_param_tensors(f"stage3 step bs6 [{tag} exp_avg]", list(net.named_parameters()),
                                   lambda p: o.state[p]["exp_avg"],
                                   lambda n: golden_util.grad_sample(ref.m[n], 65536) if ref.m[n].numel() > 65536
                                   else ref.m[n],
                                   lambda n: ref.m[n].double().norm().item(),
                                   lambda t: golden_util.grad_sample(t, 65536))


def test_fused_fc1_update_equals_unfused(sds):
    from lr2ppo_b200 import ppo
    g = torch.Generator().manual_seed(6)
    bs = 4
    text = torch.randn(bs, 2, 196, 768, generator=g).cuda()
    img = torch.randn(bs, 1, 16, 768, generator=g).repeat(1, 2, 1, 1).cuda()
    tgts = torch.randint(0, 3, (bs, 2), generator=g).cuda()
    results = []
    for fused in (True, False):
        model, reward = _build_gpu(sds)
        hp = _hp(fused)
        opt, copt, _, _ = ppo.build_optimizer(hp, model)
        mem = ppo.rollout(model, reward, text, img, tgts)
        for _ in range(2):
            ppo.update_batch(hp, model, opt, copt, mem)
        w = model.actor.out_layer.fc1.weight
        assert (w.grad is None) == fused                                  # fused: no gradient tensor exists
        results.append((golden_util.grad_sample(w.detach(), 1 << 18).cpu(),
                        golden_util.grad_sample(opt.state[w]["exp_avg"], 1 << 18).cpu(),
                        golden_util.grad_sample(model.actor._engine.bank.get(w), 1 << 18).float().cpu(),
                        model.critic.out_layer.fc1.bias.detach().cpu().clone()))
        del model, reward, opt, copt
        torch.cuda.empty_cache()
    (wf, mf, sf, bf_), (wu, mu, su, bu) = results
    assert _rel(mf, mu) < 1e-3                  # same fp32 accumulation, different tile order only
    assert (wf - wu).abs().max().item() < 1e-6
    assert torch.equal(sf, wf.to(torch.bfloat16).float())                # bf16 shadow refreshed by the fused kernel
    assert _rel(bf_, bu) < 1e-5


def test_cuda_graph_step_equals_eager(sds):
    """GraphedStage3Step (CUDA-graph replay, persistent grads, device seeds) == eager rollout + update_batch."""
    from lr2ppo_b200 import ppo
    g = torch.Generator().manual_seed(8)
    bs = 4
    batches = [(torch.randn(bs, 2, 196, 768, generator=g).cuda(),
                torch.randn(bs, 1, 16, 768, generator=g).repeat(1, 2, 1, 1).cuda(),
                torch.randint(0, 3, (bs, 2), generator=g).cuda()) for _ in range(2)]

    def build():
        model, reward = _build_gpu(sds)
        for m in model.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0                                   # train == eval, so both paths are deterministic
        hp = _hp(False)
        hp.learning_rate, hp.critic_learning_rate = 1e-7, 2e-7   # a few raw-Adam steps must stay in the linear regime
        opt, copt, _, _ = ppo.build_optimizer(hp, model)
        return model, reward, hp, opt, copt

    # eager: 2 steps on batch 0 (the graph class runs 2 eager warm-ups; the capture pass itself executes nothing)
    model, reward, hp, opt, copt = build()
    seq = [batches[0]] * 2 + [batches[1], batches[0]]
    for b in seq:
        mem = ppo.rollout(model, reward, *b)
        model.train(); st_e = ppo.update_batch(hp, model, opt, copt, mem); model.eval()
    we = golden_util.grad_sample(model.actor.out_layer.fc1.weight.detach(), 1 << 16).cpu()
    ce = model.critic.head.weight.detach().cpu().clone()
    st_e = st_e.cpu()
    del model, reward, opt, copt
    torch.cuda.empty_cache()

    model, reward, hp, opt, copt = build()
    gstep = ppo.GraphedStage3Step(hp, model, reward, opt, copt, *batches[0], warmup=2)
    gstep(*batches[1])
    st_g = gstep(*batches[0]).cpu()
    wg = golden_util.grad_sample(model.actor.out_layer.fc1.weight.detach(), 1 << 16).cpu()
    cg = model.critic.head.weight.detach().cpu()
    assert torch.allclose(st_g, st_e, rtol=1e-4, atol=1e-6), (st_g, st_e)
    assert torch.allclose(wg, we, rtol=0, atol=1e-7) and torch.allclose(cg, ce, rtol=0, atol=1e-7)


def test_graphed_cycle_equals_the_eager_cycle(sds):
    """ppo.GraphedCycle (what scripts/ppo.py runs: rollout graph + update graph over a bf16 RolloutMemory, schedulers
    stepped once per cycle, learning rate fed to the replayed AdamW through update_hyper) == the reference's cycle
    written eagerly (finetune/ppo.py:845-908: rollouts into a Python list, then train_model over it)."""
    from lr2ppo_b200 import ppo
    g = torch.Generator().manual_seed(9)
    bs, n_batches, cycles = 4, 4, 2
    data = [[(torch.randn(bs, 2, 196, 768, generator=g).cuda(), torch.randn(bs, 16, 768, generator=g).cuda(),
              torch.randint(0, 3, (bs, 2), generator=g).cuda()) for _ in range(n_batches)] for _ in range(cycles)]

    def build():
        model, reward = _build_gpu(sds)
        for m in model.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
        hp = _hp(False, bf16_grad=True)
        hp.learning_rate, hp.critic_learning_rate, hp.scheduler, hp.train_steps = 1e-7, 2e-7, "linear", 4
        return (model, reward, hp) + tuple(ppo.build_optimizer(hp, model))

    def snapshot(model, stats):
        return (golden_util.grad_sample(model.actor.out_layer.fc1.weight.detach(), 1 << 16).cpu(),
                model.critic.head.weight.detach().cpu().clone(), model.actor.text_proj.fc2.weight.detach().cpu().clone(),
                torch.stack([s.detach().float().cpu() for s in stats]))

    model, reward, hp, opt, copt, sch, csch = build()
    for cyc in data:
        mems = [ppo.rollout(model, reward, t, i.unsqueeze(1), y) for t, i, y in cyc]
        model.train()
        stats = ppo.train_model(hp, model, opt, copt, sch, csch, mems, 0)
        model.eval()
    lr_e = opt.param_groups[0]["lr"]
    eager = snapshot(model, stats)
    del model, reward, opt, copt, mems
    torch.cuda.empty_cache()

    model, reward, hp, opt, copt, sch, csch = build()
    cycle = ppo.GraphedCycle(hp, model, reward, opt, copt, capacity=n_batches, bs=bs, tags=2)
    for cyc in data:
        for t, i, y in cyc:
            cycle.rollout(t, i, y)
        assert len(cycle.memory) == n_batches
        stats = cycle.update(sch, csch)
        assert len(cycle.memory) == 0
    assert cycle._update_graph is not None and cycle._rollout_graph is not None      # the graphs really replayed
    assert opt.param_groups[0]["lr"] == lr_e
    graphed = snapshot(model, stats)
    assert torch.allclose(graphed[3], eager[3], rtol=1e-4, atol=1e-6), (graphed[3], eager[3])
    for a, b in zip(graphed[:3], eager[:3]):
        assert torch.allclose(a, b, rtol=0, atol=1e-7)


def test_two_branch_step_equals_the_single_stream_step(sds, monkeypatch):
    """LR2_DUAL_STREAM (default on: the critic's forward / backward / AdamW run on a second stream, ppo._branch_stream)
    only re-orders independent work: rollout results, the ten statistics and the updated weights are bit-identical to
    the single-stream order."""
    from lr2ppo_b200 import ppo
    g = torch.Generator().manual_seed(11)
    bs = 4
    batch = (torch.randn(bs, 2, 196, 768, generator=g).cuda(), torch.randn(bs, 1, 16, 768, generator=g).cuda(),
             torch.randint(0, 3, (bs, 2), generator=g).cuda())
    results = []
    for dual in ("1", "0"):
        monkeypatch.setenv("LR2_DUAL_STREAM", dual)
        assert (ppo._branch_stream(batch[0].device) is not None) == (dual == "1")
        model, reward = _build_gpu(sds)
        for e in (model.actor._engine, model.critic._engine):
            e.dropout_seed = 4711                    # same masks in both runs (train-mode update)
        hp = _hp(False, bf16_grad=True)
        opt, copt, _, _ = ppo.build_optimizer(hp, model)
        for _ in range(2):
            mem = ppo.rollout(model, reward, *batch)
            model.train(); stats = ppo.update_batch(hp, model, opt, copt, mem); model.eval()
        torch.cuda.synchronize()
        results.append(([t.detach().cpu().clone() for t in mem[:5]], stats.cpu(),
                        golden_util.grad_sample(model.actor.out_layer.fc1.weight.detach(), 1 << 16).cpu(),
                        golden_util.grad_sample(model.critic.out_layer.fc1.weight.detach(), 1 << 16).cpu(),
                        dict(model.critic.named_parameters())["xitt.0.0.0.fn.1.queries.weight"].detach().cpu().clone()))
        del model, reward, opt, copt
        torch.cuda.empty_cache()
    (mem_a, st_a, wa_a, wc_a, q_a), (mem_b, st_b, wa_b, wc_b, q_b) = results
    for a, b in zip(mem_a, mem_b):
        assert torch.equal(a, b)
    assert torch.equal(st_a, st_b)
    assert torch.equal(wa_a, wa_b) and torch.equal(wc_a, wc_b)
    assert torch.equal(q_a, q_b)
