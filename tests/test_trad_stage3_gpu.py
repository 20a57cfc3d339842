"""BASELINE configs[0]: one full stage-3 step (rollout + update) of the MSLR / "trad" pipeline on the CUDA path vs the
golden produced by the reference's OWN finetune/ppo_trad.py modules, `build_optimizer` and `train_model`
(oracle/make_golden.py trad_stage3).  bf16 compute: 2e-2 of scale on scores / values / rewards / statistics,
6e-2 on the norms of Adam's first moments (the gradients of a 24-row batch through two attention blocks in bf16;
lr = 0 on the first step of the linear schedule)."""
import argparse
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from lr2ppo_b200 import ppo, trad
from tests import golden_util, parity

GOLD = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "trad_stage3.pt"))


def test_trad_stage3_step_vs_reference_train_model():
    margs = argparse.Namespace(mode="reg", labels_num=5)
    model = trad.ActorCritic(margs, margs)
    model.actor.load_state_dict(golden_util.make_trad_state_dict("actor"), strict=True)
    model.critic.load_state_dict(golden_util.make_trad_state_dict("critic"), strict=True)
    reward_model = trad.Reward(margs, margs)
    reward_model.load_state_dict(golden_util.make_trad_state_dict("reward"), strict=True)
    model.cuda(); reward_model.cuda()
    for m in list(model.modules()) + list(reward_model.modules()):
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0                                   # the golden was generated without dropout noise
    text, tgts, _ = golden_util.trad_inputs("actor")
    text, tgts = text.cuda(), tgts.cuda()
    args = argparse.Namespace(is_master=False, mode="reg", kl_div_loss_weight=0.001, entropy_weight=0.001,
                              value_clip=0.5, learning_rate=1e-3, critic_learning_rate=1e-3, optimizer="adamw",
                              scheduler="linear", train_steps=100, warmup=0.1)
    opt, copt, sch, csch = ppo.build_optimizer(args, model)
    mem = trad.rollout(model, reward_model, text, tgts)
    ro = GOLD["rollout"]
    state, next_state, scores, rewards, value = mem[:5]
    assert torch.equal(next_state.cpu(), ro["next_state"])                        # permutation: bit-exact
    for got, key in ((scores, "action_scores"), (value, "value"), (rewards, "rewards")):
        ref = ro[key]
        parity.check("trad stage3", key, (got.float().cpu() - ref).abs().max().item() / max(ref.abs().max().item(), 0.1), 2e-2)
    w_before = [p.detach().clone() for p in model.parameters()]
    model.train()
    stats = trad.train_model(args, model, opt, copt, sch, csch, [mem], 0)
    got = [float(s) for s in stats]
    for i, (g, r) in enumerate(zip(got, GOLD["stats"])):
        parity.check("trad stage3", f"stat[{i}]", abs(g - r) / max(abs(r), 0.1), 2e-2)
    for p, w in zip(model.parameters(), w_before):                                 # lr = 0 on the first step
        assert torch.equal(p.detach(), w)
    checked = 0
    for tag, module, o in (("actor", model.actor, opt), ("critic", model.critic, copt)):
        for name, p in module.named_parameters():
            ref_n = GOLD[f"m/{tag}.{name}"].item()
            m1 = o.state_for(p)["exp_avg"]
            n = m1.double().norm().item()
            if ref_n < 1e-9:
                assert n < 1e-6, (tag, name, n)
                continue
            crit = tag == "critic"
            parity.check("trad stage3 [exp_avg]", f"{tag}.{name} [norm]", abs(n - ref_n) / ref_n,
                         parity._tol_for(name, parity.NORM_TOL, parity.CRITIC_TAIL_NORM if crit else None))
            full = GOLD.get(f"mfull/{tag}.{name}")
            if full is not None and ref_n > 1e-7:
                parity.check("trad stage3 [exp_avg]", f"{tag}.{name} [elem]", parity.rel_err(m1, full),
                             parity._tol_for(name, parity.GRAD_ELEM_TOL, parity.CRITIC_TAIL_ELEM if crit else None))
            checked += 1
    assert checked > 40
