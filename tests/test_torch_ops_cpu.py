"""The torch custom-op layer loads without a GPU, registers every op of its list for the CUDA key only, and refuses
CPU tensors (no fallback)."""
import pytest
import torch


def test_custom_ops_are_registered_and_cuda_only():
    from lr2ppo_b200 import torch_ops
    ns = torch_ops.load()
    for name in torch_ops.OPS:
        op = getattr(ns, name)
        schema = op.default._schema
        assert schema.name == f"lr2ppo::{name}"
    with pytest.raises(NotImplementedError):
        ns.bias_gelu(torch.zeros(2, 8), torch.zeros(8))
    with pytest.raises(NotImplementedError):
        ns.gemm(torch.zeros(4, 8, dtype=torch.bfloat16), torch.zeros(4, 8, dtype=torch.bfloat16))
    s = str(ns.gemm.default._schema)
    assert "Tensor? bias=None" in s and "int epilogue=0" in s
