"""Host model of the warp-per-query NDCG kernel (tests/ndcg_warp_model.py mirrors ndcg.cu's warp_sort<E> and the
histogram / cut / term phases index for index) against the C oracle: the register/shuffle network sorts, the padded
shared-memory layout is conflict-free, and every phase reproduces the oracle bit for bit — including heavy ties,
ragged lengths, unsorted / oversized cuts and labels outside [0, 62] (second sort of the label keys)."""
import numpy as np

from oracle import restate
from tests import ndcg_warp_model as M


def test_network_sorts_and_banks_are_distinct():
    M.main()


def _check(scores, labels, ks, lens=None):
    B, N = scores.shape
    ref, ref_order = restate.ndcg_at_k(scores, labels, ks, lens=lens, want_order=True)
    tab = restate.log2_table(N)
    for q in range(B):
        n = N if lens is None else min(int(lens[q]), N)
        out, order = M.warp_query(scores[q], labels[q], n, ks, tab)
        assert np.array_equal(order, ref_order[q, :n]), q
        assert out.tobytes() == ref[q].tobytes(), (q, out, ref[q])


def test_model_matches_oracle_histogram_path():
    rng = np.random.default_rng(1)
    ks = [1, 3, 5, 10, 20, 100000000]
    for N in (1, 2, 20, 33, 64, 100):
        _check(rng.standard_normal((4, N)).astype(np.float32), rng.integers(0, 5, (4, N)), ks)
    # heavy ties + ragged + unsorted cuts with duplicates and a zero
    N = 40
    scores = rng.integers(0, 4, (6, N)).astype(np.float32)
    scores[0, :5] = [0.0, -0.0, 0.0, -0.0, 1.0]
    labels = rng.integers(0, 63, (6, N))
    lens = rng.integers(1, N + 1, 6).astype(np.int32)
    _check(scores, labels, [7, 1, 100000000, 7, 0, 3], lens)


def test_model_matches_oracle_general_labels():
    rng = np.random.default_rng(2)
    N = 48
    scores = rng.standard_normal((5, N)).astype(np.float32)
    labels = rng.integers(0, 4, (5, N))
    labels[0, 3] = 63
    labels[1, 7] = -2
    labels[2, 0] = 70
    labels[3, 5] = 2**40
    _check(scores, labels, [1, 5, 100000000])
