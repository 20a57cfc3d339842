"""tencentpretrain/utils/misc.py: `pooling` (imported by the stage scripts, used by the classification target only)."""
import torch


def pooling(memory_bank, seg, pooling_type):
    """memory_bank [B, S, H], seg [B, S] (0 = padding) -> [B, H]."""
    m = seg.unsqueeze(-1).type_as(memory_bank)
    if pooling_type == "mean":
        return (memory_bank * m).sum(dim=1) / m.sum(dim=1)
    if pooling_type == "last":
        last = m.sum(dim=1).squeeze(-1).long() - 1
        return memory_bank[torch.arange(memory_bank.shape[0], device=memory_bank.device), last]
    if pooling_type == "max":
        return memory_bank.masked_fill(m == 0, float("-inf")).max(dim=1)[0]
    return memory_bank[:, 0]
