"""tencentpretrain/utils: the registries the stage scripts use through `from tencentpretrain.utils import *`."""
from lr2ppo_b200.optim import (AdamW, get_constant_schedule, get_constant_schedule_with_warmup,  # noqa: F401
                               get_linear_schedule_with_warmup, str2optimizer, str2scheduler)
from lr2ppo_b200.tokenizers import (BPETokenizer, CharTokenizer, SpaceTokenizer, VirtualTokenizer,  # noqa: F401
                                    str2tokenizer)
from tencentpretrain.utils.act_fun import gelu, gelu_fast, linear, relu, silu, str2act  # noqa: F401

__all__ = ["CharTokenizer", "SpaceTokenizer", "BPETokenizer", "VirtualTokenizer", "str2tokenizer", "gelu",
           "gelu_fast", "relu", "silu", "linear", "str2act", "AdamW", "str2optimizer",
           "get_linear_schedule_with_warmup", "get_constant_schedule", "get_constant_schedule_with_warmup",
           "str2scheduler"]
