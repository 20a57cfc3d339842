"""Index check of the linear-pass fused wgrad + AdamW kernel (lr2ppo_b200/csrc/adamw_wgrad.cu) on the host.

The kernel loads mma.sync.m16n8k16 fragments straight from global memory.  This model fills the fragment registers
with the kernel's own index formulas (ld_kpair rows / columns, zero rows beyond K), applies the PTX ISA fragment
layout of m16n8k16 (.row A, .col B, fp32 C/D) to multiply them, scatters the accumulators with the kernel's output
offsets, and requires the result to equal dY^T X exactly (small integers: fp32-exact) for every element of a
[out_f, in_f] weight — every CTA, warp, group and lane, including a K that is not a multiple of 16."""
import numpy as np

WG_WARPS, WG_WCOLS, WG_ROWS, WG_GROUP = 8, 128, 16, 4
WG_TCOLS = WG_WARPS * WG_WCOLS


def mma_16816(a_regs, b_regs):
    """a_regs [32 lanes, 4 regs, 2 halves], b_regs [32, 2, 2] -> c [32, 4] per the PTX fragment layout."""
    A = np.zeros((16, 16)); B = np.zeros((16, 8))
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        for i in range(8):                       # a0..a7: reg i//2, half i%2
            row = g if (i < 2 or 4 <= i < 6) else g + 8
            col = t * 2 + (i & 1) + (8 if i >= 4 else 0)
            A[row, col] = a_regs[lane, i // 2, i % 2]
        for i in range(4):                       # b0..b3
            row = t * 2 + (i & 1) + (8 if i >= 2 else 0)
            B[row, g] = b_regs[lane, i // 2, i % 2]
    D = A @ B
    c = np.zeros((32, 4))
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        for i in range(4):
            c[lane, i] = D[g + (8 if i >= 2 else 0), t * 2 + (i & 1)]
    return c


def ld_kpair(src, k, K, col):
    lo = src[k, col] if k < K else 0.0
    hi = src[k + 1, col] if k + 1 < K else 0.0
    return lo, hi


def kernel_model(dY, X, out_f, in_f):
    K = dY.shape[0]
    KS = (K + 15) // 16
    G = np.full((out_f, in_f), np.nan)
    col_tiles = (in_f + WG_TCOLS - 1) // WG_TCOLS
    for block in range((out_f // WG_ROWS) * col_tiles):
        rt, ct = divmod(block, col_tiles)
        r0 = rt * WG_ROWS
        for warp in range(WG_WARPS):
            cw = ct * WG_TCOLS + warp * WG_WCOLS
            if cw >= in_f:
                continue
            a = np.zeros((KS, 32, 4, 2))
            for ks in range(KS):
                for lane in range(32):
                    g, t = lane >> 2, lane & 3
                    k0 = ks * 16 + 2 * t
                    a[ks, lane, 0] = ld_kpair(dY, k0, K, r0 + g)
                    a[ks, lane, 1] = ld_kpair(dY, k0, K, r0 + g + 8)
                    a[ks, lane, 2] = ld_kpair(dY, k0 + 8, K, r0 + g)
                    a[ks, lane, 3] = ld_kpair(dY, k0 + 8, K, r0 + g + 8)
            for grp in range(WG_WCOLS // (8 * WG_GROUP)):
                c0 = cw + grp * 8 * WG_GROUP
                for j in range(WG_GROUP):
                    acc = np.zeros((32, 4))
                    for ks in range(KS):
                        b = np.zeros((32, 2, 2))
                        for lane in range(32):
                            g, t = lane >> 2, lane & 3
                            k0 = ks * 16 + 2 * t
                            b[lane, 0] = ld_kpair(X, k0, K, c0 + 8 * j + g)
                            b[lane, 1] = ld_kpair(X, k0 + 8, K, c0 + 8 * j + g)
                        acc += mma_16816(a[ks], b)
                    for lane in range(32):
                        g, t = lane >> 2, lane & 3
                        row_lo = (r0 + g) * in_f + c0 + 2 * t
                        row_hi = row_lo + 8 * in_f
                        for h in range(2):
                            off = (row_hi if h else row_lo) + 8 * j
                            for i in range(2):
                                r, c = divmod(off + i, in_f)
                                assert np.isnan(G[r, c]), "element written twice"
                                G[r, c] = acc[lane, 2 * h + i]
    return G


def _check(K, out_f, in_f, seed):
    rng = np.random.default_rng(seed)
    dY = rng.integers(-3, 4, (K, out_f)).astype(np.float64)
    X = rng.integers(-3, 4, (K, in_f)).astype(np.float64)
    G = kernel_model(dY, X, out_f, in_f)
    assert not np.isnan(G).any(), "element never written"
    assert np.array_equal(G, dY.T @ X)


def test_fragment_indices_k48():
    _check(48, 32, 1152, 0)          # two row tiles; 1 full CTA column tile + a partial one (1 warp)


def test_fragment_indices_ragged_k():
    _check(40, 16, 256, 1)           # K not a multiple of 16: rows 40..47 read as zero
