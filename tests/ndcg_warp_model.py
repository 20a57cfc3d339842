"""CPU emulation of the warp-per-query NDCG sort network (lr2ppo_b200/csrc/ndcg.cu: warp_sort<E>) — checks the index
math of the mirror-form bitonic network (blocked positions p = lane*E + e, register strides vs lane-xor exchanges)
and the padded shared-memory index used for the blocked payload store.  Pure numpy; no GPU."""
import numpy as np


def warp_sort(k):                       # k: [32 lanes, E] uint64, returns sorted in blocked order
    k = k.copy()
    E = k.shape[1]
    lanes = np.arange(32)

    def cex(a, b):                      # columns a < b of k: a <- min, b <- max
        lo, hi = np.minimum(k[:, a], k[:, b]), np.maximum(k[:, a], k[:, b])
        k[:, a], k[:, b] = lo, hi

    kk = 2
    while kk <= E:
        for e in range(E):
            p = e ^ (kk - 1)
            if p > e:
                cex(e, p)
        j = kk >> 2
        while j > 0:
            for e in range(E):
                if (e & j) == 0:
                    cex(e, e | j)
            j >>= 1
        kk <<= 1

    def pick(mine, other, lower):
        other_smaller = other < mine
        return np.where(other_smaller == lower, other, mine)

    m = 2
    while m <= 32:
        mask = m - 1
        lower = (lanes & (m >> 1)) == 0
        if E == 1:
            o = k[lanes ^ mask, 0]
            k[:, 0] = pick(k[:, 0], o, lower)
        else:
            for e in range(E // 2):
                o1 = k[lanes ^ mask, E - 1 - e].copy()
                o2 = k[lanes ^ mask, e].copy()
                k[:, e] = pick(k[:, e], o1, lower)
                k[:, E - 1 - e] = pick(k[:, E - 1 - e], o2, lower)
        jl = m >> 2
        while jl > 0:
            lower = (lanes & jl) == 0
            for e in range(E):
                o = k[lanes ^ jl, e].copy()
                k[:, e] = pick(k[:, e], o, lower)
            jl >>= 1
        j = E // 2
        while j > 0:
            for e in range(E):
                if (e & j) == 0:
                    cex(e, e | j)
            j >>= 1
        m <<= 1
    return k


def main():
    rng = np.random.default_rng(0)
    for E in (1, 2, 4, 8, 16, 32):
        for trial in range(20):
            n = 32 * E
            if trial % 3 == 0:
                vals = rng.integers(0, 2**63, n, dtype=np.uint64)
            elif trial % 3 == 1:
                vals = rng.integers(0, 7, n).astype(np.uint64)            # heavy ties
            else:
                vals = rng.permutation(n).astype(np.uint64)
            k = np.empty((32, E), dtype=np.uint64)
            for i, v in enumerate(vals):                                   # striped load: element i -> lane i%32, reg i//32
                k[i % 32, i // 32] = v
            out = warp_sort(k).reshape(-1)                                 # blocked: position lane*E + e
            assert np.array_equal(out, np.sort(vals)), (E, trial)
        # bank check of the padded blocked store and the striped read
        for e in range(E):
            p = np.arange(32) * E + e
            assert len(set((p + (p >> 5)) % 32)) == 32, ("store bank conflict", E, e)
            i = e * 32 + np.arange(32)
            assert len(set((i + (i >> 5)) % 32)) == 32, ("load bank conflict", E, e)
        print("E", E, "ok")
    print("NDCG_NETWORK_CHECK PASS")


if __name__ == "__main__":
    main()


# ---- the rest of ndcg_warp_kernel<E>, lane by lane (same phases, same index math) ----------------------------------
def _gain(rel):
    if rel < 0 or rel >= 64:
        g = -1
    else:
        g = (1 << int(rel)) - 1
        if g >= 2**63:
            g -= 2**64
    return np.float32(np.int64(g))


def score_u32(s):
    s = np.float32(s) + np.float32(0.0)
    u = int(np.array([s], dtype=np.float32).view(np.uint32)[0])
    u = (~u & 0xFFFFFFFF) if (u & 0x80000000) else (u | 0x80000000)
    return ~u & 0xFFFFFFFF


def label_key(l):
    asc = (int(l) ^ 0x8000000000000000) & 0xFFFFFFFFFFFFFFFF
    return ~asc & 0xFFFFFFFFFFFFFFFF


def label_from_key(k):
    v = ((~int(k)) & 0xFFFFFFFFFFFFFFFF) ^ 0x8000000000000000
    return v - 2**64 if v >= 2**63 else v


def warp_query(scores, labels, n, ks, tab):
    """One query through the warp kernel's phases.  Returns (ndcg [nk] f32, order [n])."""
    N = len(scores)
    E = 1
    while 32 * E < N:
        E *= 2
    nk = len(ks)
    hist8 = np.zeros(64 * 32, dtype=np.uint8)
    k = np.full((32, E), 0xFFFFFFFFFFFFFFFF, dtype=np.uint64)
    oor = False
    for e in range(E):
        for lane in range(32):
            i = e * 32 + lane
            if i < n:
                lab = int(labels[i])
                in_range = 0 <= lab <= 62
                lb = lab if in_range else 0xFF
                oor |= not in_range
                k[lane, e] = (score_u32(scores[i]) << 32) | (i << 8) | lb
                if in_range:
                    hist8[lb * 32 + lane] += 1
    fallback = oor
    hist = np.zeros(64, dtype=np.int64)
    for lane in range(32):
        hist[lane] = hist8[lane * 32:(lane + 1) * 32].astype(np.int64).sum()
        hist[lane + 32] = hist8[(lane + 32) * 32:(lane + 33) * 32].astype(np.int64).sum()
    hstart = np.zeros(64, dtype=np.int64)
    tot = [0] * 32
    for lane in range(32):
        d0, d1 = 2 * lane, 2 * lane + 1
        tot[lane] = (hist[62 - d0] if d0 <= 62 else 0) + (hist[62 - d1] if d1 <= 62 else 0)
    incl = np.cumsum(tot)
    for lane in range(32):
        d0, d1 = 2 * lane, 2 * lane + 1
        h0 = hist[62 - d0] if d0 <= 62 else 0
        excl = incl[lane] - tot[lane]
        if d0 <= 62:
            hstart[62 - d0] = excl
        if d1 <= 62:
            hstart[62 - d1] = excl + h0
    mine = [0] * 32
    for lane in range(nk):
        c = min(int(ks[lane]), n)
        mine[lane] = max(c, 0)
    cut_pos, cut_slot = [0] * 32, [0] * 32
    for lane in range(nk):
        rank = sum(1 for j in range(nk) if mine[j] < mine[lane] or (mine[j] == mine[lane] and j < lane))
        cut_pos[rank], cut_slot[rank] = mine[lane], lane
    k = warp_sort(k)
    pay = np.zeros(33 * E, dtype=np.uint32)
    for lane in range(32):
        for e in range(E):
            p = lane * E + e
            pay[p + (p >> 5)] = int(k[lane, e]) & 0xFFFFFFFF
    tp = np.zeros(32 * E, dtype=np.float32)
    ti = np.zeros(33 * E, dtype=np.float32)
    order = np.full(n, -1, dtype=np.int64)
    for i in range(n):
        w = int(pay[i + (i >> 5)])
        idx = w >> 8
        order[i] = idx
        lab = int(labels[idx]) if fallback else (w & 0xFF)
        tp[i] = np.float32(_gain(lab) / tab[i])
    if not fallback:
        present = [L for L in range(64) if hist[L] > 0]
        for L in sorted(present, reverse=True):
            s0, c = int(hstart[L]), int(hist[L])
            for i in range(s0, s0 + c):
                ti[i] = np.float32(_gain(L) / tab[i])
    else:
        k2 = np.full((32, E), 0xFFFFFFFFFFFFFFFF, dtype=np.uint64)
        for e in range(E):
            for lane in range(32):
                i = e * 32 + lane
                if i < n:
                    k2[lane, e] = label_key(labels[i])
        k2 = warp_sort(k2)
        for lane in range(32):
            for e in range(E):
                p = lane * E + e
                if p < n:
                    ti[p] = np.float32(_gain(label_from_key(k2[lane, e])) / tab[p])
    cut_p, cut_i = np.zeros(32, np.float32), np.zeros(32, np.float32)
    for t, outc in ((tp, cut_p), (ti, cut_i)):
        acc, i = np.float32(0), 0
        for j in range(nk):
            c = cut_pos[j]
            while i < c:
                acc = np.float32(acc + t[i])
                i += 1
            outc[cut_slot[j]] = acc
    out = np.empty(nk, dtype=np.float32)
    for j in range(nk):
        out[j] = np.float32(1.0) if cut_i[j] <= np.float32(1e-6) else np.float32(cut_p[j] / cut_i[j])
    return out, order
