"""Launch a fixed list of GEMM cases once each (for `ncu -k regex:gemm -s N -c N`).  Run on the GPU box.
   PROF_CASES=0,1 selects a subset."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lr2ppo_b200 import ops

bf = torch.bfloat16
dev = "cuda"
rnd = lambda *s: (torch.randn(*s, device=dev) * 0.05).to(bf)
Mt, E, H, items, K1 = 9408, 768, 3072, 48, 162816
x = rnd(Mt, E); w1 = rnd(H, E); b1 = torch.randn(H, device=dev); h = rnd(Mt, H); w2 = rnd(E, H)
b2 = torch.randn(E, device=dev); pre = torch.empty(Mt, H, dtype=bf, device=dev)
cat = rnd(items, K1); dy1 = rnd(items, H); gWb = torch.empty(H, K1, dtype=bf, device=dev)
out1 = torch.empty(Mt, H, dtype=bf, device=dev); out2 = torch.empty(Mt, E, dtype=bf, device=dev)
cases = [
    lambda: ops.gemm(dy1, cat, a_mn=True, b_mn=True, out=gWb, block_n=128),            # fc1 wgrad, epilogue-bound
    lambda: ops.gemm(x, w1, out=out1, block_n=2256),                                   # pair kernel, plain
    lambda: ops.gemm(x, w1, out=out1, epilogue=ops.EPI_BIAS_GELU, bias=b1, c2=pre, block_n=2256),
    lambda: ops.gemm(h, w2, out=out2, epilogue=ops.EPI_BIAS, bias=b2, block_n=2256),   # K = 3072
]
sel = os.environ.get("PROF_CASES")
if sel:
    cases = [cases[int(i)] for i in sel.split(",")]
for rep in range(2):          # pass 0 = warm-up, pass 1 = profiled (ncu -s len(cases))
    for c in cases:
        c()
    torch.cuda.synchronize()
