"""Synthetic workloads of the BASELINE configs other than the stage-3 step (which bench.py owns), shared by bench.py
(extra keys of its JSON line) and the stand-alone tools:

  configs[1]  encoder_bench()   ViT-B/16 + RoBERTa-base towers fwd + bwd (tencentpretrain build_model API)
  configs[2]  Stage2Step        stage-2 pairwise reward-model training step (reward_pair_dataloader.sh), data-parallel
  configs[4]  ndcg_point()      NDCG@k device time at one (N, B) point of the sweep

Nothing here touches oracle/ (only bench.py's baseline legs may)."""
import argparse
import os

import torch

from lr2ppo_b200 import _lib

VIT = dict(emb_size=768, feedforward_size=3072, hidden_size=768, hidden_act="gelu", heads_num=12, layers_num=12,
           dropout=0.1, max_seq_length=197, embedding=["patch", "pos"], remove_embedding_layernorm=True,
           encoder="transformer", mask="fully_visible", layernorm_positioning="pre", image_height=224,
           image_width=224, patch_size=16)                       # models/vit/base-16-224_config.json
ROBERTA = dict(emb_size=768, feedforward_size=3072, hidden_size=768, hidden_act="gelu", heads_num=12, layers_num=12,
               max_seq_length=514, dropout=0.1, embedding=["word", "pos", "seg"], encoder="transformer",
               mask="fully_visible")                              # models/xlm-roberta/base_config.json
VOCAB = 50265
VIT_GF, ROB_GF = 35.13, 11.02                                     # GFLOP per image / per 64-token sequence, forward
NDCG_KS = [1, 3, 5, 10, 20, 100000000]


def _graph(step, warm=3):
    """Capture `step` (fixed shapes, persistent buffers) in a CUDA graph after `warm` eager runs on a side stream."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(warm):
            step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        step()
    return graph


def time_fn(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


# ---------------------------------------------------------------------------------------------- configs[1] ----
def build_tower(kind):
    from lr2ppo_b200 import tower
    m = tower.build_model(argparse.Namespace(**(VIT if kind == "vit" else ROBERTA)), vocab_size=VOCAB)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "gamma" not in n and "beta" not in n:
                p.normal_(0, 0.02)
    return m.cuda()


def encoder_bench(clips=16, iters=5, train=True, graph=True, peaks=None):
    """-> {"vit": {...}, "roberta": {...}, "clips_per_s"}: model FLOP/s (backward = 2x forward) against
    the measured burst and sustained cuBLAS bf16 peaks."""
    peaks = peaks or {}
    sustained = peaks.get("bf16_tflops_sustained", 1400.0)
    burst = peaks.get("bf16_tflops", 1650.0)
    out = {"workload": f"{clips} clips x 8 keyframes 224x224 (ViT-B/16) + {clips} x 20 tags x 64 tokens (RoBERTa-base), "
                       f"fwd+bwd, dropout {'on' if train else 'off'}, bf16"}
    for kind in ("vit", "roberta"):
        m = build_tower(kind)
        m.train(train)
        if kind == "vit":
            n = clips * 8
            src = torch.randn(n, 3, 224, 224, device="cuda")
            seg = torch.ones(n, 197, dtype=torch.long, device="cuda")
            gflop, rows = 3 * VIT_GF * n, n * 197
        else:
            n = clips * 20
            src = torch.randint(5, VOCAB, (n, 64), device="cuda")
            seg = torch.ones(n, 64, dtype=torch.long, device="cuda")
            gflop, rows = 3 * ROB_GF * n, n * 64
        g = torch.randn(rows, 768, device="cuda")

        def step():
            for p in m.parameters():
                p.grad = None
            h = m(src, None, seg).float()
            (h.reshape(rows, 768) * g).sum().backward()

        c0 = _lib.launch_count()
        step()
        launches = _lib.launch_count() - c0
        run, mode = step, "eager"
        if graph:
            run, mode = _graph(step).replay, "cuda_graph_replay"
        ms = time_fn(run, iters)
        tf = gflop / ms            # GFLOP / ms = TFLOP/s
        out[kind] = {"rows": rows, "ms_fwd_bwd": round(ms, 3), "model_tflops": round(tf, 1),
                     "tensor_frac_of_sustained_peak": round(tf / sustained, 3),
                     "tensor_frac_of_burst_peak": round(tf / burst, 3), "launch_mode": mode,
                     "launches_per_step": launches, "items_per_s": round(n / ms * 1e3, 1)}
        del m, src, seg, g, run
        torch.cuda.empty_cache()
    out["clips_per_s"] = round(clips / (out["vit"]["ms_fwd_bwd"] + out["roberta"]["ms_fwd_bwd"]) * 1e3, 2)
    return out


# ---------------------------------------------------------------------------------------------- configs[2] ----
class Stage2Step:
    """One stage-2 training step (finetune/reward_pair_dataloader.py:347-365) on a synthetic batch of `pairs` label
    pairs per GPU: two forwards of the 526 M-parameter reward model over 4-slot orderings (2 x pairs*4 items), hinge
    loss, one backward, AdamW, scheduler.  world > 1: dist.GradSync averages the gradients (all-gathered fc1 wgrad
    operands + flat all-reduce of the rest)."""

    def __init__(self, dev, pairs=64, world=1, rank=0, seed=7):
        from lr2ppo_b200 import models, stages
        from lr2ppo_b200.dist import GradSync
        self.stages, self.pairs, self.world = stages, pairs, world
        margs = argparse.Namespace(mode="reg", labels_num=3, seq_length=196, max_imgs=16, visual_feat_dim=768)
        torch.manual_seed(seed)
        with torch.device(dev):
            self.model = models.PairClassifier(margs, margs)
        with torch.no_grad():
            for n, p in self.model.named_parameters():
                if "gamma" not in n and "beta" not in n:
                    p.normal_(0, 0.02)
            for mod in self.model.modules():
                if isinstance(mod, torch.nn.LayerNorm):
                    mod.weight.fill_(1.0)
        self.model.train()
        self.hp = argparse.Namespace(learning_rate=2e-5, optimizer="adamw", scheduler="linear", warmup=0.1,
                                     train_steps=68260 * 10 // pairs + 1, mode="reg",   # reward_pair_dataloader.sh
                                     fc1_grad_bf16=True, fc1_passes=2)
        self.opt, self.sch = stages.build_optimizer(self.hp, self.model)
        self.sync = None
        if world > 1:
            self.sync = GradSync(world)
            self.sync.broadcast_params(self.model)
            self.sync.attach(self.model, self.opt)
        g = torch.Generator().manual_seed(300 + rank)
        self.text = torch.randn(pairs, 2, 196, 768, generator=g).to(dev)
        self.img = torch.randn(pairs, 1, 16, 768, generator=g).repeat(1, 2, 1, 1).to(dev)
        self.tgts = torch.randint(0, 3, (pairs, 2), generator=g).to(dev)
        flip = torch.rand(pairs, generator=g) < 0.5                       # reward_pair_dataloader.py:128-141
        self.chosen = torch.tensor([[0, 1, 0, 1], [1, 0, 0, 1]])[flip.long()].to(dev)
        self.reject = torch.tensor([[0, 1, 1, 0], [1, 0, 1, 0]])[flip.long()].to(dev)

    def __call__(self):
        return self.stages.reward_train_model(self.hp, self.model, self.opt, self.sch, self.text, self.img, self.tgts,
                                              self.chosen, self.reject, grad_sync=self.sync)


# ---------------------------------------------------------------------------------------------- configs[4] ----
def ndcg_point(N, B, iters=20, labels_hi=3, seed=0):
    """Device seconds per lr2_ndcg_at_k launch on scores ~ N(0,1) [B,N], labels ~ U{0..labels_hi-1}: `iters` launches are
    captured in one CUDA graph and the replay is timed, so the ~15 us Python / ctypes cost of issuing a few-us kernel
    is not attributed to it.  Returns (seconds, algorithmic bytes = B * (N * 12 + 24))."""
    from lr2ppo_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(seed)
    scores = torch.randn(B, N, generator=g, device="cuda")
    labels = torch.randint(0, labels_hi, (B, N), generator=g, device="cuda")

    def step():
        for _ in range(iters):
            ops.ndcg_at_k(scores, labels, NDCG_KS)

    graph = _graph(step, warm=2)
    return time_fn(graph.replay, 3, warm=2) / iters * 1e-3, B * (N * 12 + 4 * len(NDCG_KS))
