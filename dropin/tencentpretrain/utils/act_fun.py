"""Activation registry (host-side definitions; inside the towers the activations are GEMM epilogues)."""
import math

import torch
import torch.nn.functional as F


def gelu(x):
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


def gelu_fast(x):
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * x ** 3)))


def relu(x):
    return F.relu(x)


def silu(x):
    return F.silu(x)


def linear(x):
    return x


str2act = {"gelu": gelu, "gelu_fast": gelu_fast, "relu": relu, "silu": silu, "linear": linear}
