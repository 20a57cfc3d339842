"""Probe: out_layer.fc1 weight gradient + AdamW in L2-sized PANELS.

Today one step of a 500 M-parameter out_layer.fc1 is two launches: wgrad writes a 1 GB bf16 gradient to HBM, AdamW reads it
back (28 B/param + 2 B shadow).  Here the [3072, 162816] matrix is cut into R x C panels; for each panel the gradient
block is written by the wgrad GEMM into ONE reused buffer and consumed by an AdamW launch over exactly that block right
after, so the gradient can live in L2 (written and re-read while hot, overwritten by the next panel before it is ever
evicted).  Prints device time per full update for several panel shapes; R = 3072, C = all = today's path.

    python tools/panel_probe.py            # on a B200
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from lr2ppo_b200 import _lib, ops

H, K1, ITEMS = 3072, 162816, 48
CHUNK = 4096


def tables(p, m, v, sh, buf, r0, r1, c0, c1, dev):
    """ptrs / meta / chunks of one AdamW launch over rows [r0, r1) x columns [c0, c1) with the gradient read from `buf`
    ([R, K1] bf16, row r of the panel = row r0 + r of the matrix)."""
    gfake = buf.data_ptr() - r0 * K1 * 2
    ptrs = torch.tensor([p.data_ptr(), gfake, m.data_ptr(), v.data_ptr(), sh.data_ptr(), 0], dtype=torch.int64, device=dev)
    wd = torch.tensor([0.01], dtype=torch.float32).view(torch.int32).item()
    meta = torch.tensor([p.numel(), wd, 1, 0], dtype=torch.int64, device=dev)
    seg = torch.arange(0, c1 - c0, CHUNK, dtype=torch.int64)
    seg_len = torch.clamp((c1 - c0) - seg, max=CHUNK)
    offs = (torch.arange(r0, r1, dtype=torch.int64)[:, None] * K1 + c0 + seg[None, :]).reshape(-1)
    ln = seg_len[None, :].expand(r1 - r0, -1).reshape(-1)
    chunks = torch.stack([ln << 32, offs], dim=1).contiguous().to(dev)
    return ptrs, meta, chunks


def main():
    dev = torch.device("cuda")
    L = _lib.load()
    g = torch.Generator(device=dev).manual_seed(0)
    p0 = torch.randn(H, K1, device=dev, generator=g) * 0.02
    dy = (torch.randn(ITEMS, H, device=dev, generator=g) * 0.1).to(torch.bfloat16)
    x = torch.randn(ITEMS, K1, device=dev, generator=g).to(torch.bfloat16)
    hyper = torch.tensor([1e-4, 0.9, 0.999, 1e-6, 0.1, 0.001, 1.0, 1e-4], device=dev)
    results, ref = [], None
    shapes = [(3072, 1), (1536, 1), (768, 1), (256, 1), (256, 2), (256, 4), (256, 6), (512, 4), (768, 6), (256, 12)]
    if len(sys.argv) > 1:
        shapes = [tuple(int(t) for t in s.split("x")) for s in sys.argv[1:]]
    for R, csplit in shapes:
        C = K1 // csplit
        assert H % R == 0 and K1 % csplit == 0 and C % 256 == 0, (R, csplit)
        p = p0.clone()
        m, v = torch.zeros_like(p), torch.zeros_like(p)
        sh = p.to(torch.bfloat16)
        buf = torch.empty(R, K1, device=dev, dtype=torch.bfloat16)
        plan = []
        for r0 in range(0, H, R):
            for c0 in range(0, K1, C):
                plan.append((r0, c0, tables(p, m, v, sh, buf, r0, r0 + R, c0, c0 + C, dev)))

        def step():
            for r0, c0, (ptrs, meta, chunks) in plan:
                ops.gemm(dy[:, r0:r0 + R], x[:, c0:c0 + C], a_mn=True, b_mn=True, out=buf[:, c0:c0 + C], block_n=2256)
                _lib.run(L.lr2_adamw_multi, ptrs.data_ptr(), meta.data_ptr(), chunks.data_ptr(), chunks.shape[0],
                         hyper.data_ptr(), _lib.stream())

        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            step()                                  # the ONE update whose result is compared across panel shapes
            snap = (p[::97, ::1013].clone(), m[::97, ::1013].clone(), sh[::97, ::1013].clone())
            step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            step()
        for _ in range(2):
            graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        if ref is None:
            ref = snap
        same = all(torch.equal(a, b) for a, b in zip(snap, ref))
        mb = R * C * 2 / 1e6
        print(f"panel {R:5d} x {C:6d} ({mb:7.1f} MB grad block, {len(plan):3d} panels): {ms:7.3f} ms per update  "
              f"= {28 * H * K1 / ms / 1e6:7.1f} GB/s on 28 B/param; identical to unpanelled: {same}", flush=True)
        results.append((R, C, ms))
        del p, m, v, sh, buf, plan, graph
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
