"""TEST INFRASTRUCTURE ONLY — generate tests/golden/* by RUNNING THE IMPORTED REFERENCE on CPU.

    python -m oracle.make_golden          (in the build container; needs /root/reference)

Outputs (committed):
  tests/golden/rows.json        NDCG, RankLoss, clipped value loss, hinge, SmoothL1, rollout sort, AdamW,
                                LR schedule, TencentPretrain LayerNorm — reference outputs on seeded inputs
  tests/golden/ppo_update.json  finetune/ppo.py:train_model executed on stub actor/critic modules whose
                                outputs are free parameters: pins losses, logged statistics and
                                d(loss)/d(scores), d(value_loss)/d(values)
  tests/golden/fusion.pt        reference Actor / Critic / Reward forward + selected gradients on
                                seed-generated weights and inputs (weights are re-generated from the seed
                                by tests/golden_util.py, they are not stored)
"""
import json
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402
from tests import golden_util  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def tl(t):
    return t.detach().cpu().tolist()


def gen_rows():
    ndcg = ref_loader.load("ndcg")
    ppo = ref_loader.load("ppo")
    out = {}
    g = torch.Generator().manual_seed(1234)

    # ---- NDCG through the reference's own call sequence (finetune/ppo.py:651-659) ----
    cases = []
    for n, hi in [(1, 3), (2, 3), (3, 3), (6, 3), (16, 3), (20, 3), (20, 5), (33, 5), (64, 3), (128, 5), (257, 3)]:
        for rep in range(2):
            scores = torch.randn(n, generator=g)
            gold = torch.randint(0, hi, (n,), generator=g)
            if rep == 1 and n > 3:
                gold = torch.zeros(n, dtype=torch.int64) if n == 6 else gold
            _, idx = torch.sort(scores, dim=-1, descending=True)
            gold_re = gold[idx]
            true_rel, _ = torch.sort(gold, dim=-1, descending=True)
            meter = ndcg.AverageNDCGMeter()
            vals = meter.return_ndcg_at_k(gold_re, true_rel)
            cases.append(dict(scores=tl(scores), labels=tl(gold), order=tl(idx), ks=meter.ndcg_at_k,
                              ndcg=[float(np.float32(v)) for v in tl(vals)]))
    out["ndcg"] = cases

    # ---- RankLoss (finetune/ppo.py:38-55) ----
    cases = []
    for B, n, margin in [(3, 2, 0.01), (24, 2, 0.01), (8, 5, 0.01), (5, 4, 1.0)]:
        s = torch.randn(B, n, generator=g) * (0.02 if margin < 0.1 else 1.0)
        order = torch.stack([torch.randperm(n, generator=g) for _ in range(B)])
        cases.append(dict(scores=tl(s), order=tl(order), margin=margin, loss=float(ppo.RankLoss(margin)(s, order))))
    s = torch.tensor([[3.0, 1.0], [2.0, 0.5]])
    cases.append(dict(scores=tl(s), order=[[0, 1], [0, 1]], margin=0.01,
                      loss=float(ppo.RankLoss(0.01)(s, torch.tensor([[0, 1], [0, 1]])))))  # all satisfied -> 0
    out["rank_loss"] = cases

    # ---- clipped value loss (finetune/ppo.py:494-498) ----
    cases = []
    for B, clip in [(2, 0.5), (24, 0.5), (24, 0.05), (7, 0.4)]:
        v = torch.randn(B, generator=g); r = torch.randn(B, generator=g); vo = torch.randn(B, generator=g) * 0.3
        vv = v.clone().requires_grad_(True)
        loss = ppo.clipped_value_loss(vv, r, vo, clip)
        loss.backward()
        cases.append(dict(v=tl(v), ret=tl(r), v_old=tl(vo), clip=clip, loss=float(loss), dv=tl(vv.grad)))
    out["value_loss"] = cases

    # ---- pair hinge (finetune/reward_pair_dataloader.py:355-358) ----
    cases = []
    for B, margin in [(64, 1.0), (24, 0.01), (3, 1.0)]:
        c = torch.randn(B, generator=g); r = torch.randn(B, generator=g)
        cc = c.clone().requires_grad_(True); rr = r.clone().requires_grad_(True)
        loss = torch.relu(margin - (cc - rr)).mean()   # line 355-356 with margin 1 / reward_trad.py:273
        loss.backward()
        acc = (c > r).float().mean()
        cases.append(dict(chosen=tl(c), reject=tl(r), margin=margin, loss=float(loss), acc=float(acc),
                          dchosen=tl(cc.grad), dreject=tl(rr.grad)))
    out["pair_hinge"] = cases

    # ---- SmoothL1 beta 0.3 (finetune/pointwise.py:229) ----
    cases = []
    for n in [40, 7]:
        x = torch.randn(n, generator=g); t = torch.randint(0, 3, (n,), generator=g)
        xx = x.clone().requires_grad_(True)
        loss = torch.nn.SmoothL1Loss(beta=0.3)(xx.view(-1), t.view(-1))
        loss.backward()
        cases.append(dict(logits=tl(x), tgt=tl(t), beta=0.3, loss=float(loss), dlogits=tl(xx.grad)))
    out["smooth_l1"] = cases

    # ---- rollout sort + compose (finetune/ppo.py:865-874) ----
    cases = []
    for B, n in [(24, 2), (5, 7), (3, 20)]:
        s = torch.randn(B, n, generator=g)
        state = torch.arange(n).unsqueeze(0).repeat(B, 1)
        _, idx = torch.sort(s, dim=-1, descending=True)
        ns = torch.stack([torch.index_select(state[i], 0, idx[i]) for i in range(B)])
        ns = torch.cat([torch.arange(2).unsqueeze(0).repeat(B, 1), ns], dim=1)
        cases.append(dict(scores=tl(s), next_state=tl(ns)))
    out["rollout"] = cases

    # ---- AdamW (tencentpretrain/utils/optimizers.py:305-402) + linear schedule (:62-86) ----
    opt_mod = ref_loader.load("tencentpretrain.utils.optimizers")
    cases = []
    for n, wd, lr, steps in [(1000, 0.01, 1e-3, 3), (37, 0.0, 5e-4, 2)]:
        p0 = torch.randn(n, generator=g) * 0.02
        grads = [torch.randn(n, generator=g) * 0.01 for _ in range(steps)]
        p = torch.nn.Parameter(p0.clone())
        opt = opt_mod.AdamW([{"params": [p], "weight_decay": wd}], lr=lr, correct_bias=False)
        for gr in grads:
            p.grad = gr.clone()
            opt.step()
        st = opt.state[p]
        cases.append(dict(p0=tl(p0), grads=[tl(x) for x in grads], wd=wd, lr=lr, p=tl(p), m=tl(st["exp_avg"]),
                          v=tl(st["exp_avg_sq"])))
    out["adamw"] = cases
    p = torch.nn.Parameter(torch.zeros(1))
    opt = opt_mod.AdamW([p], lr=1e-3, correct_bias=False)
    sch = opt_mod.get_linear_schedule_with_warmup(opt, 34130.1, 341301)
    lrs = [opt.param_groups[0]["lr"]]
    for _ in range(5):
        opt.step(); sch.step(); lrs.append(opt.param_groups[0]["lr"])
    out["schedule"] = dict(base_lr=1e-3, warmup=34130.1, total=341301, lrs=lrs)

    # ---- TencentPretrain LayerNorm (tencentpretrain/layers/layer_norm.py:5-21) ----
    ln_mod = ref_loader.load("tencentpretrain.layers.layer_norm")
    ln = ln_mod.LayerNorm(64, eps=1e-6)
    with torch.no_grad():
        ln.gamma.copy_(torch.randn(64, generator=g)); ln.beta.copy_(torch.randn(64, generator=g))
    x = torch.randn(5, 64, generator=g, requires_grad=True)
    y = ln(x)
    gy = torch.randn(5, 64, generator=g)
    (y * gy).sum().backward()
    out["tencent_ln"] = dict(x=tl(x), gamma=tl(ln.gamma), beta=tl(ln.beta), y=tl(y), gy=tl(gy), dx=tl(x.grad),
                             dgamma=tl(ln.gamma.grad), dbeta=tl(ln.beta.grad))
    with open(os.path.join(GOLD, "rows.json"), "w") as f:
        json.dump(out, f)
    print("rows.json written")


class _StubActor(torch.nn.Module):
    def __init__(self, scores):
        super().__init__()
        self.scores = torch.nn.Parameter(scores.clone())

    def forward(self, text, img, tgts):
        return torch.zeros(()), self.scores.view(-1)


class _StubCritic(torch.nn.Module):
    def __init__(self, values):
        super().__init__()
        self.values = torch.nn.Parameter(values.clone())

    def forward(self, text, img, tgts, state):
        return self.values * 1.0


class _StubAC(torch.nn.Module):
    def __init__(self, s, v):
        super().__init__()
        self.actor, self.critic = _StubActor(s), _StubCritic(v)


def gen_ppo_update():
    """Run the reference train_model (finetune/ppo.py:501-617) on stub networks."""
    import argparse
    import torch.distributed as dist
    ppo = ref_loader.load("ppo")
    if not dist.is_initialized():
        f = tempfile.NamedTemporaryFile(delete=False)
        dist.init_process_group("gloo", init_method=f"file://{f.name}", rank=0, world_size=1)
    g = torch.Generator().manual_seed(77)
    cases = []
    for B, n, scale in [(24, 2, 0.05), (24, 2, 1.0), (6, 2, 0.3)]:
        s_new = torch.randn(B, n, generator=g) * scale
        s_old = s_new + torch.randn(B, n, generator=g) * scale * 0.3
        reward = torch.randn(B, generator=g) * 0.5
        v_old = torch.randn(B, generator=g) * 0.5
        v_new = v_old + torch.randn(B, generator=g) * 0.7
        state = torch.arange(n).unsqueeze(0).repeat(B, 1)
        _, idx = torch.sort(s_old, dim=-1, descending=True)
        next_state = torch.cat([torch.arange(2).unsqueeze(0).repeat(B, 1), idx], dim=1)
        model = _StubAC(s_new, v_new)
        args = argparse.Namespace(is_master=False, mode="reg", kl_div_loss_weight=0.001, entropy_weight=0.001,
                                  value_clip=0.5)
        opt = torch.optim.SGD(model.actor.parameters(), lr=0.0)
        copt = torch.optim.SGD(model.critic.parameters(), lr=0.0)
        sch = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: 1.0)
        csch = torch.optim.lr_scheduler.LambdaLR(copt, lambda s: 1.0)
        dummy = torch.zeros(1)
        mem = [[state, next_state, s_old.clone(), reward.clone(), v_old.clone(), dummy, dummy, dummy]]
        stats = ppo.train_model(args, model, opt, copt, sch, csch, mem, 0)
        names = ["policy_loss", "value_loss", "kl", "old_value", "value", "rewards_ori", "rewards", "advantages",
                 "rank_loss", "entropy"]
        cases.append(dict(s_new=tl(s_new), s_old=tl(s_old), reward=tl(reward), v_old=tl(v_old), v_new=tl(v_new),
                          next_state=tl(next_state), w_kl=0.001, w_ent=0.001, value_clip=0.5,
                          stats={k: float(v) for k, v in zip(names, stats)},
                          ds=tl(model.actor.scores.grad), dv=tl(model.critic.values.grad)))
    with open(os.path.join(GOLD, "ppo_update.json"), "w") as f:
        json.dump(cases, f)
    print("ppo_update.json written")


def gen_fusion():
    """Reference Actor / Critic / Reward (finetune/ppo.py:196-350) forward + backward on seeded tensors."""
    ppo = ref_loader.load("ppo")
    cfg = golden_util.FUSION_CFG
    args = ref_loader.ppo_args(seq_length=cfg["seq_length"], max_imgs=cfg["max_imgs"])
    out = {"cfg": cfg}
    for kind in ("actor", "critic", "reward"):
        cls = {"actor": ppo.Actor, "critic": ppo.Critic, "reward": ppo.Reward}[kind]
        model = cls(args, args)
        golden_util.init_params_(model, golden_util.SEEDS[kind])
        model.eval()
        text, img, tgts, index = golden_util.make_inputs(kind)
        if kind == "actor":
            _, logits = model(text, img, tgts)
        else:
            logits = model(text, img, tgts, index)
        gw = golden_util.out_grad(kind, logits.numel())
        (logits * gw).sum().backward()
        rec = {"logits": logits.detach().clone()}
        for name, p in model.named_parameters():
            gr = p.grad
            rec["gnorm/" + name] = gr.double().norm().float()
            if gr.numel() <= 4096:
                rec["grad/" + name] = gr.clone()
            else:
                rec["grad/" + name] = golden_util.grad_sample(gr).clone()
        out[kind] = rec
        print(kind, "logits", logits[:4].tolist())
        del model
    torch.save(out, os.path.join(GOLD, "fusion.pt"))
    print("fusion.pt written")


def gen_trad():
    """Reference trad Actor / Critic / Reward (finetune/ppo_trad.py:142-283) forward + backward."""
    pt = ref_loader.load("ppo_trad")
    import argparse
    args = argparse.Namespace(mode="reg", labels_num=5)
    out = {}
    for kind in ("actor", "critic", "reward"):
        cls = {"actor": pt.Actor, "critic": pt.Critic, "reward": pt.Reward}[kind]
        model = cls(args, args)
        model.load_state_dict(golden_util.make_trad_state_dict(kind), strict=True)
        model.eval()
        text, tgts, index = golden_util.trad_inputs(kind)
        if kind == "actor":
            _, logits = model(text, None, tgts)
        else:
            logits = model(text, None, tgts, index)
        gw = golden_util.out_grad(kind, logits.numel())
        (logits * gw).sum().backward()
        rec = {"logits": logits.detach().clone()}
        for name, p in model.named_parameters():
            rec["gnorm/" + name] = p.grad.double().norm().float()
            rec["grad/" + name] = (p.grad if p.grad.numel() <= 4096 else golden_util.grad_sample(p.grad)).clone()
        out[kind] = rec
        print("trad", kind, logits[:3].tolist())
    torch.save(out, os.path.join(GOLD, "trad.pt"))


def gen_stage12():
    """ONE reference training step of stage 1 (finetune/pointwise.py:300-313) and stage 2
    (finetune/reward_pair_dataloader.py:347-365) with the reference AdamW + constant schedule."""
    import argparse
    opt_mod = ref_loader.load("tencentpretrain.utils.optimizers")
    out = {}
    for stage in (1, 2):
        mod = ref_loader.load("pointwise" if stage == 1 else "reward_pair_dataloader")
        cfg = golden_util.FUSION_CFG
        args = argparse.Namespace(mode="reg", labels_num=3, seq_length=cfg["seq_length"], max_imgs=cfg["max_imgs"],
                                  visual_feat_dim=cfg["feat"])
        model = mod.Classifier(args, args)
        kind = "actor" if stage == 1 else "reward"
        model.load_state_dict(golden_util.make_state_dict(kind), strict=True)
        model.train()
        for m in model.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0                      # parity is defined without dropout noise
        no_decay = ["bias", "gamma", "beta"]
        named = list(model.named_parameters())
        groups = [{"params": [p for n, p in named if not any(nd in n for nd in no_decay)], "weight_decay": 0.01},
                  {"params": [p for n, p in named if any(nd in n for nd in no_decay)], "weight_decay": 0.0}]
        opt = opt_mod.AdamW(groups, lr=golden_util.STEP_LR, correct_bias=False)
        sch = opt_mod.get_constant_schedule(opt)
        text, img, tgts, chosen, reject = golden_util.stage_inputs(stage)
        before = {n: p.detach().clone() for n, p in named}
        if stage == 1:
            loss = mod.train_model(args, model, opt, sch, text, img, tgts)
            rec = {"loss": loss.detach().clone()}
        else:
            loss, acc = mod.train_model(args, model, opt, sch, text, img, tgts, chosen, reject)
            rec = {"loss": loss.detach().clone(), "acc": acc.detach().clone()}
        for n, p in named:
            rec["m/" + n] = golden_util.grad_sample(opt.state[p]["exp_avg"], 4096).clone()
            rec["mnorm/" + n] = opt.state[p]["exp_avg"].double().norm().float()
            rec["delta/" + n] = golden_util.grad_sample(p.detach() - before[n], 4096).clone()
        out[f"stage{stage}"] = rec
        print("stage", stage, {k: v.tolist() for k, v in rec.items() if k in ("loss", "acc")})
        del model, opt, before
    torch.save(out, os.path.join(GOLD, "stage12.pt"))


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    which = sys.argv[1:] or ["rows", "ppo_update", "fusion", "trad", "stage12"]
    if "rows" in which:
        gen_rows()
    if "ppo_update" in which:
        gen_ppo_update()
    if "fusion" in which:
        gen_fusion()
    if "trad" in which:
        gen_trad()
    if "stage12" in which:
        gen_stage12()


def _tower_args(kind):
    """argparse namespace the reference's build_model needs: opts defaults < JSON config (utils/config.py:6-23)."""
    import argparse
    import json as js
    opts = ref_loader.load("tencentpretrain.opts")
    parser = argparse.ArgumentParser()
    opts.model_opts(parser)
    args = parser.parse_args([])
    cfg = js.load(open(os.path.join(ref_loader.REF, "models", "vit/base-16-224_config.json" if kind == "vit"
                                    else "xlm-roberta/base_config.json")))
    for k, v in cfg.items():
        setattr(args, k, v)
    args.tokenizer = argparse.Namespace(vocab=list(range(golden_util.TOWER_VOCAB)))
    args.tie_weights = False
    args.has_lmtarget_bias = False
    args.labels_num = 3
    return args


def gen_tower():
    """Reference build_model (tencentpretrain/model_builder.py:8) ViT-B/16 and RoBERTa-base: embedding + encoder
    forward and backward on seeded inputs."""
    mb = ref_loader.load("tencentpretrain.model_builder")
    out = {}
    for kind in ("vit", "roberta"):
        args = _tower_args(kind)
        model = mb.build_model(args)
        model.eval()
        names = [(n, tuple(p.shape)) for n, p in model.named_parameters() if not n.startswith("target.")]
        sd = golden_util.make_tower_state_dict(names, golden_util.TOWER_SEEDS[kind])
        missing = model.load_state_dict(sd, strict=False)
        assert all(k.startswith("target.") for k in missing.missing_keys), missing
        src, seg = golden_util.tower_inputs(kind)
        hidden = model.encoder(model.embedding(src, seg), seg)
        gw = golden_util.out_grad("actor", hidden.numel()).view_as(hidden)
        (hidden * gw).sum().backward()
        rec = {"names": names, "hidden": hidden.detach().clone()}
        for n, p in model.named_parameters():
            if n.startswith("target.") or p.grad is None:
                continue
            rec["gnorm/" + n] = p.grad.double().norm().float()
            rec["grad/" + n] = (p.grad if p.grad.numel() <= 4096 else golden_util.grad_sample(p.grad)).clone()
        out[kind] = rec
        print("tower", kind, hidden.shape, hidden.flatten()[:3].tolist())
    torch.save(out, os.path.join(GOLD, "tower.pt"))


if __name__ == "__main__" and "tower" in sys.argv[1:]:
    gen_tower()


def gen_api():
    """Reference VideoTransformer (finetune/video_transformer.py:8) and ProjectionLayer
    (finetune/project_embedding.py:5), never built by the scripts: API-parity goldens."""
    vt = ref_loader.load("video_transformer")
    pe = ref_loader.load("project_embedding")
    out = {}
    for kind in ("video", "proj"):
        model = vt.VideoTransformer(**golden_util.VIDEO_CFG) if kind == "video" else pe.ProjectionLayer(768, 512, 0.2)
        model.eval()
        names = [(n, tuple(p.shape)) for n, p in model.named_parameters()]
        model.load_state_dict(golden_util.make_api_state_dict(names, golden_util.API_SEEDS[kind]), strict=True)
        x = golden_util.api_input(kind).requires_grad_(True)
        y = model(x)
        gw = golden_util.out_grad("critic", y.numel()).view_as(y)
        (y * gw).sum().backward()
        rec = {"names": names, "y": y.detach().clone(), "dx": x.grad.clone()}
        for n, p in model.named_parameters():
            rec["gnorm/" + n] = p.grad.double().norm().float()
            rec["grad/" + n] = (p.grad if p.grad.numel() <= 4096 else golden_util.grad_sample(p.grad)).clone()
        out[kind] = rec
        print("api", kind, y.shape, y.flatten()[:3].tolist())
    torch.save(out, os.path.join(GOLD, "api.pt"))


if __name__ == "__main__" and "api" in sys.argv[1:]:
    gen_api()


def gen_trad_stage3():
    """BASELINE configs[0]: one full stage-3 step of finetune/ppo_trad.py on the reference's own trad Actor / Critic /
    Reward modules and its own `train_model` / `build_optimizer` (CPU, gloo world 1).  The rollout is inline in the
    reference's main() (ppo_trad.py:765-813), so those lines are re-executed here against the reference modules.
    Dropout is switched off (p = 0) so the step is deterministic; the linear schedule starts at lr = 0, so the first
    step moves no weight and the gradients are pinned through Adam's first moments."""
    import argparse
    import torch.distributed as dist
    pt = ref_loader.load("ppo_trad")
    if not dist.is_initialized():
        f = tempfile.NamedTemporaryFile(delete=False)
        dist.init_process_group("gloo", init_method=f"file://{f.name}", rank=0, world_size=1)
    margs = argparse.Namespace(mode="reg", labels_num=5)
    model = pt.ActorCritic(margs, margs)
    model.actor.load_state_dict(golden_util.make_trad_state_dict("actor"), strict=True)
    model.critic.load_state_dict(golden_util.make_trad_state_dict("critic"), strict=True)
    reward_model = pt.Reward(margs, margs)
    reward_model.load_state_dict(golden_util.make_trad_state_dict("reward"), strict=True)
    for m in list(model.modules()) + list(reward_model.modules()):
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    text, tgts, _ = golden_util.trad_inputs("actor")
    args = argparse.Namespace(is_master=False, mode="reg", kl_div_loss_weight=0.001, entropy_weight=0.001,
                              value_clip=0.5, learning_rate=1e-3, critic_learning_rate=1e-3, optimizer="adamw",
                              scheduler="linear", train_steps=100, warmup=0.1, device=torch.device("cpu"))
    opt, copt, sch, csch = pt.build_optimizer(args, model)
    bs, tags_num = text.shape[:2]
    # ---- rollout, ppo_trad.py:765-813 ----
    model.eval(); reward_model.eval()
    state = torch.tensor([i for i in range(tags_num)]).unsqueeze(0).repeat(bs, 1)
    with torch.no_grad():
        _, action_logits = model.actor(text, None, tgts)
        value = model.critic(text, None, tgts, state)
    action_scores = action_logits.view(bs, tags_num)
    _, idx = torch.sort(action_scores, dim=-1, descending=True)
    next_state = torch.stack([torch.index_select(state[i], 0, idx[i]) for i in range(bs)])
    next_state = torch.cat([torch.arange(2).unsqueeze(0).repeat(bs, 1), next_state], dim=1)
    with torch.no_grad():
        rewards = reward_model(text, None, tgts, next_state)
    memories = [[state.clone().detach(), next_state.clone().detach(), action_scores.clone().detach(),
                 rewards.clone().detach(), value.clone().detach(), text.clone().detach(), tgts.clone().detach()]]
    out = {"rollout": dict(action_scores=action_scores.clone(), value=value.clone(), next_state=next_state.clone(),
                           rewards=rewards.clone())}
    model.train()
    stats = pt.train_model(args, model, opt, copt, sch, csch, memories, 0)
    out["stats"] = [float(s) for s in stats]
    for tag, module, o in (("actor", model.actor, opt), ("critic", model.critic, copt)):
        for name, p in module.named_parameters():
            m1 = o.state[p]["exp_avg"]
            out[f"m/{tag}.{name}"] = m1.double().norm().float()
            if m1.numel() <= 4096:
                out[f"mfull/{tag}.{name}"] = m1.clone()
    torch.save(out, os.path.join(GOLD, "trad_stage3.pt"))
    print("trad_stage3.pt written; stats", out["stats"])


if __name__ == "__main__" and "trad_stage3" in sys.argv[1:]:
    gen_trad_stage3()


def gen_ppo_helpers():
    """The PaLM-rlhf-style helpers the reference keeps next to its update (finetune/ppo.py:431-491: log, log_prob,
    masked_entropy, masked_kl_div, masked_normalize) run on seeded inputs: they pin the pieces from which the
    north_star's ratio-clipped surrogate and sampled-ranking log-probability are restated (oracle/restate.py), since the
    reference has no implementation of either.  Also pins a torch.autograd value of the clipped surrogate composed
    from those helpers exactly as PaLM-rlhf composes them."""
    ppo = ref_loader.load("ppo")
    g = torch.Generator().manual_seed(4321)
    out = {"normalize": [], "log_prob": [], "kl": [], "entropy": [], "surrogate": []}
    for B in (1, 2, 24, 200):
        t = torch.randn(B, generator=g) * 3 + 1
        out["normalize"].append(dict(t=tl(t), out=tl(ppo.masked_normalize(t))))
    out["normalize"].append(dict(t=[0.5, 0.5, 0.5], out=tl(ppo.masked_normalize(torch.tensor([0.5, 0.5, 0.5])))))
    for B, n in ((3, 2), (5, 7)):
        s = torch.randn(B, n, generator=g)
        p = torch.softmax(s, -1)
        q = torch.softmax(torch.randn(B, n, generator=g), -1)
        out["log_prob"].append(dict(p=tl(p), out=tl(ppo.log_prob(p))))
        out["kl"].append(dict(p=tl(p), q=tl(q), out=tl(ppo.masked_kl_div(p, q))))
        out["entropy"].append(dict(p=tl(p), out=tl(ppo.masked_entropy(p))))
    for B, eps, normalize in ((24, 0.2, True), (24, 0.2, False), (7, 0.05, True), (1, 0.2, False)):
        logp = (torch.randn(B, generator=g) * 0.3).requires_grad_(True)
        logp_old = torch.randn(B, generator=g) * 0.3
        adv = torch.randn(B, generator=g)
        a = ppo.masked_normalize(adv) if normalize else adv
        ratios = (logp - logp_old).exp()
        surr1 = ratios * a
        surr2 = ratios.clamp(1 - eps, 1 + eps) * a
        loss = (-torch.min(surr1, surr2)).mean()
        loss.backward()
        out["surrogate"].append(dict(logp=tl(logp), logp_old=tl(logp_old), adv=tl(adv), eps=eps, normalize=normalize,
                                     loss=float(loss), dlogp=tl(logp.grad)))
    with open(os.path.join(GOLD, "ppo_helpers.json"), "w") as f:
        json.dump(out, f)
    print("ppo_helpers.json written")


if __name__ == "__main__" and "ppo_helpers" in sys.argv[1:]:
    gen_ppo_helpers()
