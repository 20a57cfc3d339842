"""Device time of the tower attention kernels at the BASELINE configs[1] shapes (ViT-B/16: B = 128, S = 197;
RoBERTa-base: B = 320, S = 64), dropout on / off, CUDA-graph replay of 10 launches.  Run on the GPU box."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from lr2ppo_b200 import ops


def timeit(fn, iters=10):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 3 / iters * 1e3


def main():
    H, E = 12, 768
    for name, B, S in (("ViT-B/16", 128, 197), ("RoBERTa-base", 320, 64)):
        qkv = (torch.randn(B * S, 3 * E, device="cuda") * 0.5).to(torch.bfloat16)
        d_o = torch.randn(B * S, E, device="cuda").to(torch.bfloat16)
        bias = torch.zeros(B, S, device="cuda")
        flop_f = 4.0 * S * S * 64 * B * H
        for p in (0.1, 0.0):
            o, lse = ops.mha_fwd(qkv, B, S, H, key_bias=bias, drop_p=p, seed=1)
            uf = timeit(lambda: ops.mha_fwd(qkv, B, S, H, key_bias=bias, drop_p=p, seed=1))
            ub = timeit(lambda: ops.mha_bwd(qkv, o, d_o, lse, B, S, H, key_bias=bias, drop_p=p, seed=1))
            print(f"{name:13s} B={B:4d} S={S:4d} dropout={p}: fwd {uf:7.1f} us ({flop_f / uf / 1e6:6.1f} TFLOP/s)   "
                  f"bwd {ub:7.1f} us ({2.5 * flop_f / ub / 1e6:6.1f} TFLOP/s)", flush=True)


if __name__ == "__main__":
    main()
