"""Round-2 validation + timing of the linear-pass fused wgrad+AdamW kernel (adamw_wgrad.cu, LR2_WGRAD_ADAMW_IMPL=mma).
Run on a B200:   LR2_WGRAD_ADAMW_IMPL=mma python tools/check_wgrad_mma.py          (then without the variable: tcgen05)
1. small shape vs a torch fp32 restatement of the update (tolerance 1e-5 of scale: fp32 accumulate both sides);
2. full out_layer.fc1 shape [3072, 162816], K = 48: device time per launch and effective TB/s on 26 B/element.
Then:  LR2_WGRAD_ADAMW_IMPL=mma python -m pytest tests/test_stage3_gpu.py -k fused_fc1 -q"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lr2ppo_b200 import _lib

HYPER = (1e-3, 0.9, 0.999, 1e-6, 0.1, 0.001, 1.0, 1e-3)     # step_size, b1, b2, eps, 1-b1, 1-b2, grad_scale, lr


def launch(L, dy, x, p, m, v, sh, hyper, wd):
    _lib.run(L.lr2_gemm_wgrad_adamw, dy.data_ptr(), dy.stride(0), x.data_ptr(), x.stride(0), dy.shape[0], p.shape[0],
             p.shape[1], p.data_ptr(), m.data_ptr(), v.data_ptr(), sh.data_ptr(), hyper.data_ptr(), float(wd),
             _lib.stream())


def main():
    L = _lib.load()
    dev = "cuda"
    hyper = torch.tensor(HYPER, dtype=torch.float32, device=dev)
    g = torch.Generator(device=dev).manual_seed(0)
    impl = os.environ.get("LR2_WGRAD_ADAMW_IMPL", "tcgen05")
    shapes = ((48, 64, 1152), (40, 32, 256), (16, 16, 128)) if impl == "mma" else ((48, 256, 1280), (40, 128, 256))
    for K, out_f, in_f in shapes:      # (the tcgen05 implementation works on 128-row x 128/256-column tiles)
        dy = (torch.randn(K, out_f, generator=g, device=dev) * 0.1).bfloat16()
        x = torch.randn(K, in_f, generator=g, device=dev).bfloat16()
        p = torch.randn(out_f, in_f, generator=g, device=dev)
        m = torch.randn(out_f, in_f, generator=g, device=dev) * 0.01
        v = torch.rand(out_f, in_f, generator=g, device=dev) * 0.01
        sh = torch.zeros(out_f, in_f, dtype=torch.bfloat16, device=dev)
        p0, m0, v0 = p.clone(), m.clone(), v.clone()
        launch(L, dy, x, p, m, v, sh, hyper, 0.01)
        torch.cuda.synchronize()
        gr = dy.float().t() @ x.float()
        lr, b1, b2, eps, omb1, omb2, gs, lrd = HYPER
        gr = gr * gs
        mr = m0 * b1 + gr * omb1
        vr = v0 * b2 + gr * gr * omb2
        pr = p0 - lr * (mr / (vr.sqrt() + eps))
        pr = pr - lrd * 0.01 * pr
        for name, a, b in (("p", p, pr), ("m", m, mr), ("v", v, vr)):
            err = (a - b).abs().max().item() / b.abs().max().item()
            assert err < 1e-5, (impl, K, out_f, in_f, name, err)
        assert torch.equal(sh, p.bfloat16()), "bf16 shadow"
        print(f"[{impl}] K={K} [{out_f},{in_f}] ok")
    out_f, in_f, K = 3072, 162816, 48
    dy = (torch.randn(K, out_f, generator=g, device=dev) * 0.1).bfloat16()
    x = torch.randn(K, in_f, generator=g, device=dev).bfloat16()
    p = torch.randn(out_f, in_f, device=dev)
    m = torch.zeros_like(p); v = torch.zeros_like(p)
    sh = torch.zeros(out_f, in_f, dtype=torch.bfloat16, device=dev)
    for _ in range(2):
        launch(L, dy, x, p, m, v, sh, hyper, 0.01)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        launch(L, dy, x, p, m, v, sh, hyper, 0.01)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"[{impl}] full fc1 shape: {ms:.3f} ms/launch = {out_f * in_f * 26 / ms / 1e9:.2f} TB/s on 26 B/element "
          f"(unfused today: wgrad 0.26 ms + AdamW 2.14 ms)")


if __name__ == "__main__":
    main()
