"""One attention forward + backward (ViT-B/16 shape) for `ncu -k regex:mha_tc`.  Run on the GPU box."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lr2ppo_b200 import ops
B, S, H, E = 128, 197, 12, 768
qkv = (torch.randn(B * S, 3 * E, device="cuda") * 0.5).to(torch.bfloat16)
d_o = torch.randn(B * S, E, device="cuda").to(torch.bfloat16)
bias = torch.zeros(B, S, device="cuda")
for _ in range(2):
    o, lse = ops.mha_fwd(qkv, B, S, H, key_bias=bias, drop_p=0.1, seed=1)
    d = ops.mha_bwd(qkv, o, d_o, lse, B, S, H, key_bias=bias, drop_p=0.1, seed=1)
torch.cuda.synchronize()
