"""AverageNDCGMeter with the reference's interface (ndcg.py:9-65) on the fused NDCG kernels.

`return_ndcg_at_k(predicted_relevance, true_relevances)` takes two int64 relevance lists already in
rank order (CUDA tensors) and returns the [len(ndcg_at_k)] fp32 vector, bit-identical to the
reference's sequential fp32 arithmetic.  `batch_ndcg(scores, labels)` is the batched fast path
(segmented sort + DCG for B queries in one launch) used by evaluation.
"""
import torch

from . import ops


class AverageNDCGMeter(object):
    def __init__(self, ndcg_at_k=[1, 3, 5, 10, 20, 100000000]):
        self.ndcg = {}
        self.ndcg_at_k = ndcg_at_k
        self.reset()

    def reset(self):
        for k in self.ndcg_at_k:
            self.ndcg[k] = []

    def value(self):
        # mean over queries; like the reference this turns the lists into tensors (single use)
        for k in self.ndcg:
            self.ndcg[k] = torch.mean(torch.stack(self.ndcg[k]))
        return self.ndcg

    def _pair(self, predicted_relevance, true_relevances):
        if not predicted_relevance.is_cuda:
            raise RuntimeError("AverageNDCGMeter runs on CUDA tensors (no CPU fallback)")
        p = predicted_relevance.to(torch.int64).contiguous().view(1, -1)
        t = true_relevances.to(torch.int64).contiguous().view(1, -1)
        if p.shape != t.shape:
            raise ValueError("predicted and true relevance lists must have the same length")
        return ops.ndcg_presorted(p, t, self.ndcg_at_k)[0]

    def return_ndcg_at_k(self, predicted_relevance, true_relevances):
        return self._pair(predicted_relevance, true_relevances)

    def compute_ndcg_at_k(self, predicted_relevance, true_relevances):
        vals = self._pair(predicted_relevance, true_relevances)
        for i, k_val in enumerate(self.ndcg_at_k):
            self.ndcg[k_val].append(vals[i])

    def compute_ndcg_at_k_batch(self, predicted_relevance, true_relevances):
        assert predicted_relevance.shape == true_relevances.shape
        vals = ops.ndcg_presorted(predicted_relevance.to(torch.int64).contiguous(),
                                  true_relevances.to(torch.int64).contiguous(), self.ndcg_at_k)
        for row in vals:
            for i, k_val in enumerate(self.ndcg_at_k):
                self.ndcg[k_val].append(row[i])

    # ---- batched path (new): scores + labels for B queries, one launch -------------------
    def batch_ndcg(self, scores, labels, lens=None, want_order=False):
        """scores f32 [B,N], labels i64 [B,N] (ragged via lens i32 [B]) -> ndcg [B, len(ks)]."""
        return ops.ndcg_at_k(scores.float().contiguous(), labels.to(torch.int64).contiguous(), self.ndcg_at_k,
                             lens=lens, want_order=want_order)

    def add_batch(self, ndcg_rows):
        for row in ndcg_rows:
            for i, k_val in enumerate(self.ndcg_at_k):
                self.ndcg[k_val].append(row[i])
