"""BASELINE configs[1]: stage-1 encoder towers (ViT-B/16 + RoBERTa-base, TencentPretrain build_model API) forward +
backward on a synthetic LRMovieNet-shaped batch (B clips x 8 keyframes 224x224, 20 tags x 64 tokens per clip), bf16
compute on one B200.  Reports model FLOP/s (ViT 35.13 GFLOP/img fwd, RoBERTa 11.02 GFLOP/seq fwd at
S=64; backward = 2x forward) against the measured bf16 peaks.  Run on the GPU box:
    python tools/encoder_bench.py [clips]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from lr2ppo_b200 import tower, _lib

VIT = dict(emb_size=768, feedforward_size=3072, hidden_size=768, hidden_act="gelu", heads_num=12, layers_num=12,
           dropout=0.1, max_seq_length=197, embedding=["patch", "pos"], remove_embedding_layernorm=True,
           encoder="transformer", mask="fully_visible", layernorm_positioning="pre", image_height=224,
           image_width=224, patch_size=16)                       # models/vit/base-16-224_config.json
ROBERTA = dict(emb_size=768, feedforward_size=3072, hidden_size=768, hidden_act="gelu", heads_num=12, layers_num=12,
               max_seq_length=514, dropout=0.1, embedding=["word", "pos", "seg"], encoder="transformer",
               mask="fully_visible")                              # models/xlm-roberta/base_config.json
VOCAB = 50265
VIT_GF, ROB_GF = 35.13, 11.02


def build(kind):
    args = argparse.Namespace(**(VIT if kind == "vit" else ROBERTA))
    m = tower.build_model(args, vocab_size=VOCAB)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "gamma" not in n and "beta" not in n:
                p.normal_(0, 0.02)
    return m.cuda()


def timeit(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0 = _lib.launch_count()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, (_lib.launch_count() - c0) // iters


def main():
    clips = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    iters = int(os.environ.get("ENC_ITERS", "5"))
    train = os.environ.get("ENC_TRAIN", "1") == "1"
    torch.manual_seed(7)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                            "MEASURED_PEAKS.json")))
    except Exception:
        pass
    sustained = peaks.get("bf16_tflops_sustained", 1400.0)
    out = {"config": f"configs[1]: {clips} clips x 8 keyframes 224x224 (ViT-B/16) + {clips}x20 tags x 64 tokens "
                     f"(RoBERTa-base), fwd+bwd, dropout {'on' if train else 'off'}", "peak_tflops_sustained": sustained}
    for kind in ("vit", "roberta"):
        m = build(kind)
        m.train(train)
        if kind == "vit":
            n = clips * 8
            src = torch.randn(n, 3, 224, 224, device="cuda")
            seg = torch.ones(n, 197, dtype=torch.long, device="cuda")
            gflop = 3 * VIT_GF * n
            rows = n * 197
        else:
            n = clips * 20
            src = torch.randint(5, VOCAB, (n, 64), device="cuda")
            seg = torch.ones(n, 64, dtype=torch.long, device="cuda")
            gflop = 3 * ROB_GF * n
            rows = n * 64
        g = torch.randn(rows, 768, device="cuda")

        def step():
            for p in m.parameters():
                p.grad = None
            h = m(src, None, seg).float()
            (h.reshape(rows, 768) * g).sum().backward()

        mode = "eager"
        run = step
        if os.environ.get("ENC_GRAPH", "1") == "1":
            # the ~400 launches of a tower step are replayed as one CUDA graph (dropout seeds come from a device
            # counter bumped inside the graph, so every replay draws fresh masks)
            try:
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    for _ in range(3):
                        step()
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    step()
                run, mode = graph.replay, "cuda_graph_replay"
            except Exception as e:          # report, do not hide
                print(f"graph capture failed for {kind}: {e!r}; timing the eager step", file=sys.stderr)
                torch.cuda.synchronize()
        ms, launches = timeit(run, iters)
        if mode != "eager":
            launches = None
        if os.environ.get("ENC_PROFILE") == "1":
            total_e0, total_e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            _lib.PROFILE = []
            total_e0.record(); step(); total_e1.record()
            torch.cuda.synchronize()
            prof, _lib.PROFILE = _lib.PROFILE, None
            agg = {}
            for name, a, e0, e1 in prof:
                key = name
                if name == "lr2_gemm_bf16":
                    key = f"gemm M={a[10]} N={a[11]} K={a[12]} amn={a[2]} bmn={a[5]} epi={a[13]}"
                t = agg.setdefault(key, [0, 0.0]); t[0] += 1; t[1] += e0.elapsed_time(e1)
            tot = sum(v[1] for v in agg.values())
            print(f"-- {kind}: eager step {total_e0.elapsed_time(total_e1):.2f} ms, sum of C-ABI calls {tot:.2f} ms", file=sys.stderr)
            for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:18]:
                print(f"   {k:70s} n={v[0]:3d} {v[1]:8.3f} ms {100 * v[1] / tot:5.1f}%", file=sys.stderr)
        tf = gflop / ms            # GFLOP / ms = TFLOP/s
        out[kind] = {"rows": rows, "ms_fwd_bwd": round(ms, 3), "model_tflops": round(tf, 1),
                     "frac_of_sustained_bf16_peak": round(tf / sustained, 3), "launch_mode": mode,
                     "launches_per_step": launches,
                     "items_per_s": round(n / ms * 1e3, 1)}
        del m, src, seg, g
        torch.cuda.empty_cache()
    clip_ms = out["vit"]["ms_fwd_bwd"] + out["roberta"]["ms_fwd_bwd"]
    out["clips_per_s"] = round(clips / clip_ms * 1e3, 2)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
