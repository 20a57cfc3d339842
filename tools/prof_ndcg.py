"""ncu driver for the NDCG kernels: launch 1 warms up, launch 2 is config-5's largest point (N=1024, B=4096), launch 3 a
few long queries (N=1024, B=64; with LR2_NDCG_WARP=1 also on the warp-per-query kernel).
  LR2_NDCG_WARP=1 ncu --set full --import-source on -k regex:ndcg_warp -s 1 -c 2 -o gpurun_out/prof_ndcg python tools/prof_ndcg.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lr2ppo_b200 import ops

KS = [1, 3, 5, 10, 20, 100000000]
g = torch.Generator(device="cuda").manual_seed(0)
scores = torch.randn(4096, 1024, generator=g, device="cuda")
labels = torch.randint(0, 5, (4096, 1024), generator=g, device="cuda")
for B in (4096, 4096, 64):
    out = ops.ndcg_at_k(scores[:B], labels[:B], KS)
torch.cuda.synchronize()
print("ok", float(out.mean()))
