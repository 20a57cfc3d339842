"""torch custom-op layer (`torch.ops.lr2ppo.*`) over the C ABI -- SURVEY.md §8(b): one op per C symbol of the hot path.

    from lr2ppo_b200 import torch_ops            # loads lr2ppo_b200/liblr2ppo_torch.so (built by `make torch`)
    y = torch.ops.lr2ppo.gemm(x, w, epilogue=2, bias=b)                 # D = x w^T + b, exact-erf GELU, bf16
    ndcg, order = torch.ops.lr2ppo.ndcg_at_k(scores, labels, ks, log2_table)

The ops are registered with TORCH_LIBRARY for the CUDA dispatch key only (lr2ppo_b200/csrc/torch_binding.cpp): a CPU
tensor raises NotImplementedError, there is no fallback.  Each op validates device / dtype / contiguity, allocates its
outputs, calls the matching `lr2_*` entry point on at::cuda::getCurrentCUDAStream() and raises on a non-zero return
code; all of them can be captured in CUDA graphs.  The Python module mirror (`lr2ppo_b200.ops`, ctypes) and this layer
call the SAME entry points: tests/test_torch_ops_gpu.py checks the results bit for bit and reports the per-call host
overhead of both bindings."""
import os

import torch

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "liblr2ppo_torch.so")
OPS = ["gemm", "layernorm_fwd", "layernorm_bwd", "xit_attention_fwd", "xit_attention_bwd", "flash_attention_fwd",
       "flash_attention_bwd", "gather_items", "bias_gelu", "dropout_philox", "ppo_rollout", "ppo_policy_loss",
       "clipped_value_loss", "pair_hinge_loss", "smooth_l1", "gae_scan", "ndcg_at_k", "adamw_multi_tensor"]
_loaded = False


def load():
    """Load the custom-op library once; raises when it has not been built (there is no fallback)."""
    global _loaded
    if not _loaded:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} not found: build it with `make -C lr2ppo_b200/csrc torch` "
                               "(python -c 'import __graft_entry__ as g; g.build()' does)")
        torch.ops.load_library(LIB_PATH)
        _loaded = True
    return torch.ops.lr2ppo


load()
