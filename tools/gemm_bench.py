"""Micro-benchmark of the LR2PPO GEMM shapes (stage-3, 48 items) through the C ABI, with the library next to it:
`cuBLAS` = torch.matmul on the same bf16 operands and layouts (GEMM only), `eager` = torch.matmul + the ATen kernels
the fused epilogue replaces (bias add, exact GELU, dropout, residual, GELU backward).  Run on the GPU box."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lr2ppo_b200 import ops

bf = torch.bfloat16
dev = "cuda"


def rnd(*s):
    return (torch.randn(*s, device=dev) * 0.05).to(bf)


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us


def main():
    Mt, E, H, items, K1 = 9408, 768, 3072, 48, 162816
    x = rnd(Mt, E); w1 = rnd(H, E); b1 = torch.randn(H, device=dev); h = rnd(Mt, H); w2 = rnd(E, H)
    b2 = torch.randn(E, device=dev); pre = torch.empty(Mt, H, dtype=bf, device=dev); res = rnd(Mt, E)
    wq = rnd(E, E); dy = rnd(Mt, E); dh = rnd(Mt, H)
    W = rnd(H, K1); cat = rnd(items, K1); dy1 = rnd(items, H); bh = torch.randn(H, device=dev)
    gW = torch.empty(H, K1, dtype=torch.float32, device=dev)
    gWb = torch.empty(H, K1, dtype=bf, device=dev)
    gw1 = torch.empty(H, E, dtype=torch.float32, device=dev)
    y1 = torch.empty(items, H, dtype=bf, device=dev); dcat = torch.empty(items, K1, dtype=bf, device=dev)
    out1 = torch.empty(Mt, H, dtype=bf, device=dev); out2 = torch.empty(Mt, E, dtype=bf, device=dev)
    cases = [
        ("fwd 9408x3072x768 plain", 2 * Mt * H * E, (Mt * E + H * E + Mt * H) * 2,
         lambda: ops.gemm(x, w1, out=out1)),
        ("fwd 9408x3072x768 bias", 2 * Mt * H * E, (Mt * E + H * E + Mt * H) * 2,
         lambda: ops.gemm(x, w1, out=out1, epilogue=ops.EPI_BIAS, bias=b1)),
        ("fwd 9408x3072x768 bias+gelu+pre", 2 * Mt * H * E, (Mt * E + H * E + 2 * Mt * H) * 2,
         lambda: ops.gemm(x, w1, out=out1, epilogue=ops.EPI_BIAS_GELU, bias=b1, c2=pre)),
        ("fwd 9408x3072x768 gelu+drop", 2 * Mt * H * E, (Mt * E + H * E + 2 * Mt * H) * 2,
         lambda: ops.gemm(x, w1, out=out1, epilogue=ops.EPI_BIAS_GELU, bias=b1, c2=pre, drop_p=0.1, seed=1, site=2)),
        ("fwd 9408x768x3072 bias", 2 * Mt * H * E, (Mt * H + H * E + Mt * E) * 2,
         lambda: ops.gemm(h, w2, out=out2, epilogue=ops.EPI_BIAS, bias=b2)),
        ("fwd 9408x768x3072 bias+drop+res", 2 * Mt * H * E, (Mt * H + H * E + 2 * Mt * E) * 2,
         lambda: ops.gemm(h, w2, out=out2, epilogue=ops.EPI_BIAS_DROP_RES, bias=b2, aux=res, drop_p=0.1, seed=1, site=3)),
        ("fwd 9408x768x768 bias", 2 * Mt * E * E, (2 * Mt * E + E * E) * 2,
         lambda: ops.gemm(x, wq, out=out2, epilogue=ops.EPI_BIAS, bias=b2)),
        ("dgrad 9408x768x3072 (b_mn)", 2 * Mt * H * E, (Mt * H + H * E + Mt * E) * 2,
         lambda: ops.gemm(dh, w1, b_mn=True, out=out2)),
        ("dgrad 9408x3072x768 dgelu (b_mn)", 2 * Mt * H * E, (Mt * E + H * E + 2 * Mt * H) * 2,
         lambda: ops.gemm(dy, w2, b_mn=True, out=out1, epilogue=ops.EPI_DGELU, aux=pre)),
        ("wgrad 3072x768 K=9408 f32", 2 * Mt * H * E, (Mt * H + Mt * E) * 2 + H * E * 4,
         lambda: ops.gemm(dh, x, a_mn=True, b_mn=True, out=gw1)),
        ("fc1 fwd 3072x48x162816 s6 T", 2 * items * H * K1, (H * K1 + items * K1) * 2,
         lambda: ops.gemm(W, cat, out=y1, transposed_out=True, epilogue=ops.EPI_BIAS_GELU, bias=bh, splits=6, block_n=64)),
        ("fc1 dgrad 162816x48x3072 T", 2 * items * H * K1, (H * K1 + items * K1) * 2,
         lambda: ops.gemm(W, dy1, a_mn=True, out=dcat, transposed_out=True, block_n=64)),
        ("fc1 wgrad 3072x162816 K=48 f32", 2 * items * H * K1, H * K1 * 4 + items * K1 * 2,
         lambda: ops.gemm(dy1, cat, a_mn=True, b_mn=True, out=gW)),
        # what the training step runs: bf16 gradient buffer, pair kernel + TMA-store drain (engine._fc1_wgrad)
        ("fc1 wgrad 3072x162816 K=48 bf16 (step)", 2 * items * H * K1, H * K1 * 2 + items * K1 * 2,
         lambda: ops.gemm(dy1, cat, a_mn=True, b_mn=True, out=gWb, block_n=2256)),
    ]
    for bn in [int(v) for v in os.environ.get("BENCH_BNS", "").split(",") if v]:
        cases += [
            (f"fwd 9408x3072x768 plain BN={bn}", 2 * Mt * H * E, (Mt * E + H * E + Mt * H) * 2,
             lambda bn=bn: ops.gemm(x, w1, out=out1, block_n=bn)),
            (f"fwd 9408x3072x768 gelu+pre BN={bn}", 2 * Mt * H * E, (Mt * E + H * E + 2 * Mt * H) * 2,
             lambda bn=bn: ops.gemm(x, w1, out=out1, epilogue=ops.EPI_BIAS_GELU, bias=b1, c2=pre, block_n=bn)),
            (f"fwd 9408x3072x768 gelu+drop BN={bn}", 2 * Mt * H * E, (Mt * E + H * E + 2 * Mt * H) * 2,
             lambda bn=bn: ops.gemm(x, w1, out=out1, epilogue=ops.EPI_BIAS_GELU, bias=b1, c2=pre, drop_p=0.1, seed=1,
                                    site=2, block_n=bn)),
            (f"fwd 9408x768x3072 bias BN={bn}", 2 * Mt * H * E, (Mt * H + H * E + Mt * E) * 2,
             lambda bn=bn: ops.gemm(h, w2, out=out2, epilogue=ops.EPI_BIAS, bias=b2, block_n=bn)),
            (f"fwd 9408x768x768 bias BN={bn}", 2 * Mt * E * E, (2 * Mt * E + E * E) * 2,
             lambda bn=bn: ops.gemm(x, wq, out=out2, epilogue=ops.EPI_BIAS, bias=b2, block_n=bn)),
            (f"dgrad 9408x768x3072 (b_mn) BN={bn}", 2 * Mt * H * E, (Mt * H + H * E + Mt * E) * 2,
             lambda bn=bn: ops.gemm(dh, w1, b_mn=True, out=out2, block_n=bn)),
            (f"dgrad 9408x3072x768 dgelu (b_mn) BN={bn}", 2 * Mt * H * E, (Mt * E + H * E + 2 * Mt * H) * 2,
             lambda bn=bn: ops.gemm(dy, w2, b_mn=True, out=out1, epilogue=ops.EPI_DGELU, aux=pre, block_n=bn)),
            (f"wgrad 3072x768 K=9408 f32 BN={bn}", 2 * Mt * H * E, (Mt * H + Mt * E) * 2 + H * E * 4,
             lambda bn=bn: ops.gemm(dh, x, a_mn=True, b_mn=True, out=gw1, block_n=bn)),
            (f"wgrad 3072x768 K=9408 f32 s2 BN={bn}", 2 * Mt * H * E, (Mt * H + Mt * E) * 2 + H * E * 4,
             lambda bn=bn: ops.gemm(dh, x, a_mn=True, b_mn=True, out=gw1, block_n=bn, splits=2)),
            (f"fc1 wgrad 3072x162816 K=48 bf16 BN={bn}", 2 * items * H * K1, H * K1 * 2 + items * K1 * 2,
             lambda bn=bn: ops.gemm(dy1, cat, a_mn=True, b_mn=True, out=gWb, block_n=bn)),
        ]
    import torch.nn.functional as F
    # library equivalents by case-name prefix: (cuBLAS GEMM only, GEMM + unfused ATen epilogue)
    lib = {
        "fwd 9408x3072x768 plain": (lambda: torch.matmul(x, w1.t()), None),
        "fwd 9408x3072x768 bias": (lambda: torch.matmul(x, w1.t()), lambda: F.linear(x, w1, b1.to(bf))),
        "fwd 9408x3072x768 bias+gelu+pre": (lambda: torch.matmul(x, w1.t()), lambda: F.gelu(F.linear(x, w1, b1.to(bf)))),
        "fwd 9408x3072x768 gelu+drop": (lambda: torch.matmul(x, w1.t()),
                                        lambda: F.dropout(F.gelu(F.linear(x, w1, b1.to(bf))), 0.1, True)),
        "fwd 9408x768x3072 bias": (lambda: torch.matmul(h, w2.t()), lambda: F.linear(h, w2, b2.to(bf))),
        "fwd 9408x768x3072 bias+drop+res": (lambda: torch.matmul(h, w2.t()),
                                            lambda: F.dropout(F.linear(h, w2, b2.to(bf)), 0.1, True) + res),
        "fwd 9408x768x768 bias": (lambda: torch.matmul(x, wq.t()), lambda: F.linear(x, wq, b2.to(bf))),
        "dgrad 9408x768x3072 (b_mn)": (lambda: torch.matmul(dh, w1), None),
        "dgrad 9408x3072x768 dgelu (b_mn)": (lambda: torch.matmul(dy, w2),
                                             lambda: torch.ops.aten.gelu_backward(torch.matmul(dy, w2), pre)),
        "wgrad 3072x768 K=9408 f32": (lambda: torch.matmul(dh.t(), x), None),
        "fc1 fwd 3072x48x162816 s6 T": (lambda: torch.matmul(cat, W.t()), lambda: F.gelu(F.linear(cat, W, bh.to(bf)))),
        "fc1 dgrad 162816x48x3072 T": (lambda: torch.matmul(dy1, W), None),
        "fc1 wgrad 3072x162816 K=48 f32": (lambda: torch.matmul(dy1.t(), cat), None),
        "fc1 wgrad 3072x162816 K=48 bf16 (step)": (lambda: torch.matmul(dy1.t(), cat), None),
    }
    only = sys.argv[1] if len(sys.argv) > 1 else None
    print(f"{'case':42s} {'ours us':>9s} {'TFLOP/s':>9s} {'GB/s':>9s} | {'cuBLAS us':>9s} {'ours/cuBLAS':>11s} | "
          f"{'eager us':>9s} {'speed-up':>8s}")
    for name, flops, byts, fn in cases:
        if only and only not in name:
            continue
        us = timeit(fn)
        line = f"{name:42s} {us:9.1f} {flops / us / 1e6:9.1f} {byts / us / 1e3:9.1f}"
        cub, eag = lib.get(name, (None, None))
        if cub is not None:
            uc = timeit(cub)
            line += f" | {uc:9.1f} {uc / us:10.2f}x"
            if eag is not None:
                ue = timeit(eag)
                line += f" | {ue:9.1f} {ue / us:7.2f}x"
        print(line, flush=True)
    # AdamW on a 500M tensor
    from lr2ppo_b200.optim import FusedAdamW
    p = torch.nn.Parameter(torch.randn(H, K1, device=dev) * 0.02)
    p.grad = gW
    opt = FusedAdamW([p], lr=1e-4, weight_decay=0.01, correct_bias=False, shadow_bf16=True)
    us = timeit(lambda: opt.step(), 5)
    n = p.numel()
    print(f"{'adamw 500M fp32 grad + bf16 shadow':42s} {us:9.1f} us  {'':8s}           {30 * n / us / 1e3:8.1f} GB/s")
    from lr2ppo_b200 import _lib
    L = _lib.load()
    st = opt.state[p]
    sh = opt.shadow_of(p)
    hyper = torch.tensor([1e-4, 0.9, 0.999, 1e-6, 0.1, 0.001, 1.0, 1e-4], device=dev)
    fn = lambda: _lib.run(L.lr2_gemm_wgrad_adamw, dy1.data_ptr(), dy1.stride(0), cat.data_ptr(), cat.stride(0), items,
                          H, K1, p.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), sh.data_ptr(),
                          hyper.data_ptr(), 0.01, _lib.stream())
    us = timeit(fn, 5)
    print(f"{'fused fc1 wgrad+adamw (26 B/param)':42s} {us:9.1f} us  {'':8s}           {26 * n / us / 1e3:8.1f} GB/s")


if __name__ == "__main__":
    main()
