"""torch.ops.lr2ppo.* (TORCH_LIBRARY layer, csrc/torch_binding.cpp) against the ctypes module mirror (lr2ppo_b200.ops):
both call the same C entry points, so results must be bit-identical; CUDA-graph capture works; the per-call host
overhead of the two bindings is measured and recorded (gpurun_out/binding_overhead.json)."""
import json
import os
import time

import pytest
import torch

pytestmark = pytest.mark.gpu

from lr2ppo_b200 import ops, torch_ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
T = torch_ops.load()
bf = torch.bfloat16


def _g(seed=0):
    return torch.Generator(device="cuda").manual_seed(seed)


def test_gemm_layernorm_attention_match_the_module_mirror():
    g = _g()
    x = torch.randn(392, 768, generator=g, device="cuda").to(bf)
    w = (torch.randn(3072, 768, generator=g, device="cuda") * 0.02).to(bf)
    b = torch.randn(3072, generator=g, device="cuda") * 0.02
    pre_a = torch.empty(392, 3072, dtype=bf, device="cuda")
    pre_b = torch.empty_like(pre_a)
    assert torch.equal(T.gemm(x, w, epilogue=ops.EPI_BIAS_GELU, bias=b, c2=pre_a),
                       ops.gemm(x, w, epilogue=ops.EPI_BIAS_GELU, bias=b, c2=pre_b, splits=1))
    assert torch.equal(pre_a, pre_b)
    assert torch.equal(T.gemm(x, w, b_mn=False, out_f32=True), ops.gemm(x, w, out_dtype=torch.float32, splits=1))
    gam, bet = torch.rand(768, device="cuda") + 0.5, torch.randn(768, device="cuda")
    y1, st1 = T.layernorm_fwd(x, gam, bet, 1e-5)
    y2, st2 = ops.layernorm_fwd(x, gam, bet, 1e-5)
    assert torch.equal(y1, y2) and torch.equal(st1, st2)
    dy = torch.randn(392, 768, generator=g, device="cuda").to(bf)
    dx1, dg1, db1 = T.layernorm_bwd(dy, x, gam, st1, 1e-5)
    dx2, _, dg2, db2 = ops.layernorm_bwd(dy, x, gam, st2, 1e-5)
    assert torch.equal(dx1, dx2) and torch.equal(dg1, dg2) and torch.equal(db1, db2)
    q = torch.randn(4, 196, 768, generator=g, device="cuda").to(bf)
    k = torch.randn(4, 16, 768, generator=g, device="cuda").to(bf)
    v = torch.randn(4, 16, 768, generator=g, device="cuda").to(bf)
    post = 768 ** -0.5
    assert torch.equal(T.xit_attention_fwd(q, k, v, 8, 1.0, post), ops.xattn_fwd(q, k, v, 8, 1.0, post))
    d_o = torch.randn(4, 196, 768, generator=g, device="cuda").to(bf)
    for a, c in zip(T.xit_attention_bwd(q, k, v, d_o, 8, 1.0, post), ops.xattn_bwd(q, k, v, d_o, 8, 1.0, post)):
        assert torch.equal(a, c)
    qkv = torch.randn(2 * 197, 3 * 768, generator=g, device="cuda").to(bf)
    o1, l1 = T.flash_attention_fwd(qkv, 2, 197, 12, None, 0.125)
    o2, l2 = ops.mha_fwd(qkv, 2, 197, 12)
    assert torch.equal(o1, o2) and torch.equal(l1, l2)
    assert torch.equal(T.flash_attention_bwd(qkv, o1, o1, l1, 2, 197, 12, None, 0.125), ops.mha_bwd(qkv, o2, o2, l2, 2, 197, 12))


def test_row_kernels_ndcg_and_glue_match_the_module_mirror():
    g = _g(1)
    s = torch.randn(24, 2, generator=g, device="cuda") * 0.1
    so = s + torch.randn(24, 2, generator=g, device="cuda") * 0.05
    rew, vold = torch.randn(24, generator=g, device="cuda"), torch.randn(24, generator=g, device="cuda")
    pi = torch.stack([torch.randperm(2) for _ in range(24)]).cuda()
    scal, kl, ent, radj, adv, ds = T.ppo_policy_loss(s, so, rew, vold, pi, 0.001, 0.001)
    r = ops.ppo_policy_loss(s, so, rew, vold, pi, 0.001, 0.001)
    assert torch.equal(scal[0], r["loss"]) and torch.equal(ds, r["ds"]) and torch.equal(kl, r["kl"])
    assert torch.equal(T.ppo_rollout(s, torch.arange(2, device="cuda").repeat(24, 1), 2), ops.ppo_rollout(s, torch.arange(2, device="cuda").repeat(24, 1), 2))
    l1, dv1 = T.clipped_value_loss(rew, vold, so[:, 0].contiguous(), 0.5)
    l2, dv2 = ops.clipped_value_loss(rew, vold, so[:, 0].contiguous(), 0.5)
    assert torch.equal(l1[0], l2) and torch.equal(dv1, dv2)
    out, dc, dr = T.pair_hinge_loss(rew, vold, 1.0)
    lo, ac, dc2, dr2 = ops.pair_hinge_loss(rew, vold, 1.0)
    assert torch.equal(out[0], lo) and torch.equal(out[1], ac) and torch.equal(dc, dc2) and torch.equal(dr, dr2)
    tg = torch.randint(0, 3, (24,), device="cuda")
    a1, a2 = T.smooth_l1(rew, tg, 0.3), ops.smooth_l1_loss(rew, tg, 0.3)
    assert torch.equal(a1[0][0], a2[0]) and torch.equal(a1[1], a2[1])
    rw, val = torch.randn(5, 7, generator=g, device="cuda"), torch.randn(5, 8, generator=g, device="cuda")
    for a, c in zip(T.gae_scan(rw, val, 0.99, 0.95), ops.gae_scan(rw, val, 0.99, 0.95)):
        assert torch.equal(a, c)
    sc = torch.randn(64, 100, generator=g, device="cuda")
    lab = torch.randint(0, 3, (64, 100), device="cuda")
    ks = torch.tensor([1, 3, 5, 10, 20, 100000000], device="cuda")
    n1, o1 = T.ndcg_at_k(sc, lab, ks, ops.log2_table(100, sc.device))
    n2, o2 = ops.ndcg_at_k(sc, lab, [1, 3, 5, 10, 20, 100000000], want_order=True)
    assert torch.equal(n1, n2) and torch.equal(o1, o2)
    src = torch.randn(6, 2, 196, 768, generator=g, device="cuda")
    idx = torch.randint(0, 2, (6, 4), device="cuda")
    assert torch.equal(T.gather_items(src, idx), ops.cast_gather(src, idx))
    xf, bias = torch.randn(48, 3072, generator=g, device="cuda"), torch.randn(3072, generator=g, device="cuda")
    assert torch.equal(T.bias_gelu(xf, bias), ops.bias_gelu_rows(xf, bias, want_pre=False)[0])
    xb = torch.randn(64, 768, generator=g, device="cuda").to(bf)
    assert torch.equal(T.dropout_philox(xb, 0.1, 5, 2), ops.dropout(xb, 0.1, 5, 2))
    with pytest.raises(RuntimeError):
        T.ndcg_at_k(sc.double(), lab, ks, ops.log2_table(100, sc.device))          # wrong dtype is an error


def test_custom_ops_capture_in_a_cuda_graph_and_binding_overhead():
    g = _g(2)
    x = torch.randn(128, 768, generator=g, device="cuda").to(bf)
    w = (torch.randn(768, 768, generator=g, device="cuda") * 0.02).to(bf)
    b = torch.zeros(768, device="cuda")
    for _ in range(3):
        T.gemm(x, w, epilogue=ops.EPI_BIAS, bias=b)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        T.gemm(x, w, epilogue=ops.EPI_BIAS, bias=b)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        y = T.gemm(x, w, epilogue=ops.EPI_BIAS, bias=b)
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(y, ops.gemm(x, w, epilogue=ops.EPI_BIAS, bias=b, splits=1))
    # host time per call (the kernels are tiny and queue up: wall clock / calls = issue cost of the binding)
    gam, bet = torch.ones(768, device="cuda"), torch.zeros(768, device="cuda")
    res = {}
    for name, fn in (("torch.ops.lr2ppo.layernorm_fwd", lambda: T.layernorm_fwd(x, gam, bet, 1e-5)),
                     ("ops.layernorm_fwd (ctypes)", lambda: ops.layernorm_fwd(x, gam, bet, 1e-5)),
                     ("torch.ops.lr2ppo.gemm", lambda: T.gemm(x, w, epilogue=ops.EPI_BIAS, bias=b)),
                     ("ops.gemm (ctypes)", lambda: ops.gemm(x, w, epilogue=ops.EPI_BIAS, bias=b, splits=1))):
        for _ in range(50):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(2000):
            fn()
        dt = time.perf_counter() - t0
        torch.cuda.synchronize()
        res[name] = round(dt / 2000 * 1e6, 2)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "binding_overhead.json"), "w") as f:
        json.dump({"host_us_per_call": res}, f, indent=1)
    print("host us per call:", res)
    assert res["torch.ops.lr2ppo.layernorm_fwd"] < 40 and res["torch.ops.lr2ppo.gemm"] < 40
