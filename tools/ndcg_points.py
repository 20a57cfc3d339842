"""Device time of lr2_ndcg_at_k at the corner points of the BASELINE configs[4] sweep (no CPU oracle: bench.py
--ndcg-sweep is the full table).  LR2_NDCG_WARP64=1 selects the round-1 64-bit-key warp kernel for comparison."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools import workloads

if __name__ == "__main__":
    peak = 6536.0
    out = {}
    for N in (16, 32, 64, 128, 256, 512, 1024):
        for B in (64, 1024, 4096):
            sec, byts = workloads.ndcg_point(N, B)
            out[f"N{N}_B{B}"] = {"us": round(sec * 1e6, 2), "GB_per_s": round(byts / sec / 1e9, 1),
                                 "frac": round(byts / sec / 1e9 / peak, 4)}
    print(json.dumps({"kernel": "warp64" if os.environ.get("LR2_NDCG_WARP64") == "1" else "warp32", "points": out}))
