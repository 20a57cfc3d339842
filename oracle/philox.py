"""TEST INFRASTRUCTURE ONLY — numpy restatement of the dropout mask stream of the CUDA path
(lr2ppo_b200/csrc/common.cuh: dropout_mult8, dropout_thresh16, dropout_scale16).

The reference draws its dropout masks from torch's global RNG (finetune/xit.py:26-41, three nn.Dropout(0.1) per XiT
block); the CUDA path draws them from a counter-based generator so that backward can regenerate them: ONE
Philox4x32-7 call per aligned group of 8 elements, keyed by the 64-bit seed, counter = (group index lo, hi, site, 0x38),
each of the four output words split into two 16-bit fields; element j is kept iff its field >= round(p * 65536) and the
survivors are scaled by 65536 / (65536 - round(p * 65536)).  The mask is therefore a pure function of
(seed, site, linear element index, p): this module reproduces it on the CPU, tests/test_dropout_replay_gpu.py checks it
bit for bit against lr2_dropout_bf16, and oracle/make_golden_r2.py replays it through the reference's own modules
(nn.Dropout replaced by a multiply with these masks) to pin train-mode parity."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
LOW = np.uint64(0xFFFFFFFF)


def thresh16(p):
    t = np.float32(p) * np.float32(65536.0) + np.float32(0.5)
    return int(min(max(float(t), 0.0), 65535.0))


def scale16(p):
    return float(np.float32(65536.0) / (np.float32(65536.0) - np.float32(thresh16(p)))) if p > 0 else 1.0


def fields16(seed, site, n_groups, first_group=0):
    """uint16 [n_groups, 8]: the eight 16-bit fields of Philox4x32-7 for groups first_group .. first_group+n_groups."""
    idx = np.arange(first_group, first_group + n_groups, dtype=np.uint64)
    c0 = idx & LOW
    c1 = idx >> np.uint64(32)
    c2 = np.full(n_groups, site, dtype=np.uint64)
    c3 = np.full(n_groups, 0x38, dtype=np.uint64)
    k0 = int(seed) & 0xFFFFFFFF
    k1 = (int(seed) >> 32) & 0xFFFFFFFF
    for _ in range(7):
        p0 = M0 * c0                      # 32 x 32 -> 64-bit products (operands < 2^32, no overflow in uint64)
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & LOW
        hi1, lo1 = p1 >> np.uint64(32), p1 & LOW
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    out = np.empty((n_groups, 8), dtype=np.uint16)
    for j, c in enumerate((c0, c1, c2, c3)):
        out[:, 2 * j] = (c & np.uint64(0xFFFF)).astype(np.uint16)
        out[:, 2 * j + 1] = ((c >> np.uint64(16)) & np.uint64(0xFFFF)).astype(np.uint16)
    return out


def dropout_multiplier(seed, site, numel, p):
    """float32 [numel]: keep ? 65536/(65536 - thresh16) : 0 for linear element indices 0..numel-1 (numel % 8 == 0)."""
    assert numel % 8 == 0
    f = fields16(seed, site, numel // 8).reshape(-1)
    return np.where(f >= thresh16(p), np.float32(scale16(p)), np.float32(0.0)).astype(np.float32)


def xit_masks(seed, site_base, rows, emb, hidden, p=(0.1, 0.1, 0.1)):
    """The three masks of one XiT block evaluated on `rows` query rows, as torch tensors keyed 1, 2, 3:
    site_base+1: after the attention output projection [rows, emb]      (finetune/xit.py:35)
    site_base+2: inside the FFN after GELU             [rows, hidden]   (finetune/xit.py:109)
    site_base+3: after the FFN                          [rows, emb]      (finetune/xit.py:40)
    (lr2ppo_b200/engine.py: xit_forward uses exactly these sites and the row-major output index.)"""
    import torch
    shapes = {1: (rows, emb), 2: (rows, hidden), 3: (rows, emb)}
    return {k: torch.from_numpy(dropout_multiplier(seed, site_base + k, r * c, p[k - 1]).reshape(r, c))
            for k, (r, c) in shapes.items()}
