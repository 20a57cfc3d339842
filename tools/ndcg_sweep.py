"""BASELINE config 5: NDCG@k throughput sweep (label-set sizes 16-1024, batch 64-4096) on one B200, with the CPU
oracle (oracle/rows.c, the C restatement of ndcg.py) timed beside it.  Writes a markdown table."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from lr2ppo_b200 import ops
from oracle import restate

KS = [1, 3, 5, 10, 20, 100000000]


def gpu_time(scores, labels, iters=20):
    """Device time per call: the `iters` calls are captured in one CUDA graph and the replay is timed, so the Python /
    ctypes cost of issuing a 3 us kernel (about 15 us per call) is not attributed to the kernel."""
    for _ in range(3):
        ops.ndcg_at_k(scores, labels, KS)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        ops.ndcg_at_k(scores, labels, KS)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for _ in range(iters):
            out = ops.ndcg_at_k(scores, labels, KS)
    for _ in range(2):
        graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    graph.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/ndcg_sweep.md"
    peak = 6536.0
    try:
        peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
    except Exception:
        pass
    rows = []
    rng = np.random.default_rng(0)
    for N in (16, 32, 64, 128, 256, 512, 1024):
        for B in (64, 256, 1024, 4096):
            s = rng.standard_normal((B, N)).astype(np.float32)
            l = rng.integers(0, 3, (B, N))
            sg, lg = torch.tensor(s, device="cuda"), torch.tensor(l, device="cuda")
            t = gpu_time(sg, lg)
            ok = ops.ndcg_at_k(sg, lg, KS).cpu().numpy().tobytes() == restate.ndcg_at_k(s[:64], l[:64], KS).tobytes() \
                if B == 64 else None
            nb = min(B, 256)
            t0 = time.perf_counter()
            restate.ndcg_at_k(s[:nb], l[:nb], KS)
            tc = (time.perf_counter() - t0) / nb * B
            byts = B * (N * 12 + 4 * len(KS))
            rows.append((N, B, t * 1e6, B / t, byts / t / 1e9, byts / t / 1e9 / peak, tc * 1e6, tc / t, ok))
            print(rows[-1], flush=True)
    with open(out, "w") as f:
        f.write("# NDCG@k sweep (BASELINE config 5), 1x B200 vs CPU oracle (oracle/rows.c, 1 thread)\n\n")
        f.write(f"Algorithmic bytes = B*(N*(4+8) + 24); HBM peak = {peak} GB/s (measured copy).  GPU us = device time per "
                f"launch (20 launches replayed as one CUDA graph).\n\n")
        f.write("| N | B | GPU us | queries/s | GB/s | frac of HBM peak | CPU us (1 thread) | speed-up | bit-exact |\n")
        f.write("|---:|---:|---:|---:|---:|---:|---:|---:|:-:|\n")
        for r in rows:
            f.write(f"| {r[0]} | {r[1]} | {r[2]:.1f} | {r[3]:.3g} | {r[4]:.1f} | {r[5]:.4f} | {r[6]:.0f} | {r[7]:.0f}x | "
                    f"{'yes' if r[8] else ('' if r[8] is None else 'NO')} |\n")


if __name__ == "__main__":
    main()
