"""tcgen05 GEMM parity (floating point -> torch fp32 reference of the same op, tolerance in each test)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from lr2ppo_b200 import ops


def _rel(d, ref):
    return ((d.float() - ref).abs().max() / ref.abs().max().clamp_min(1e-6)).item()


def _mk(rows, k, mn, gen):
    t = torch.randn((k, rows) if mn else (rows, k), generator=gen, device="cuda", dtype=torch.float32)
    return t.to(torch.bfloat16)


def _ref(a, b, a_mn, b_mn):
    A = a.float().t() if a_mn else a.float()
    B = b.float().t() if b_mn else b.float()
    return A @ B.t()


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K,bn", [(128, 128, 64, 128), (256, 256, 512, 128), (200, 136, 72, 0),
                                       (384, 64, 320, 64), (304, 256, 192, 256), (1000, 768, 768, 0)])
def test_gemm_majors(M, N, K, bn, a_mn, b_mn):
    # MN-major operands need the rows dimension to be a multiple of 8 (16-byte pitch)
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    a = _mk(M, K, a_mn, g)
    b = _mk(N, K, b_mn, g)
    out = ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, out_dtype=torch.float32, block_n=bn)
    ref = _ref(a, b, a_mn, b_mn)
    torch.cuda.synchronize()
    assert _rel(out, ref) < 2e-3, f"rel err {_rel(out, ref)}"   # bf16 products are exact in fp32; only sum order differs


def test_gemm_persistent_many_tiles():
    g = torch.Generator(device="cuda").manual_seed(1)
    a = _mk(9408, 768, False, g)
    b = _mk(3072, 768, False, g)
    out = ops.gemm(a, b, out_dtype=torch.float32)
    ref = a.float() @ b.float().t()
    assert _rel(out, ref) < 2e-3


@pytest.mark.parametrize("splits", [2, 5, 16])
@pytest.mark.parametrize("transposed", [False, True])
def test_gemm_splitk_transposed(splits, transposed):
    g = torch.Generator(device="cuda").manual_seed(splits)
    M, N, K = 384, 48, 4096
    a = _mk(M, K, False, g)
    b = _mk(N, K, False, g)
    bias_len = M if transposed else N
    bias = torch.randn(bias_len, generator=g, device="cuda")
    if transposed:
        out = ops.gemm(a, b, transposed_out=True, splits=splits, out_dtype=torch.float32, epilogue=ops.EPI_BIAS,
                       bias=bias, out=torch.empty((N, M), device="cuda"))
        ref = (a.float() @ b.float().t()).t() + bias[None, :]
    else:
        out = ops.gemm(a, b, splits=splits, out_dtype=torch.float32, epilogue=ops.EPI_BIAS, bias=bias)
        ref = a.float() @ b.float().t() + bias[None, :]
    assert _rel(out, ref) < 2e-3


def test_gemm_bias_gelu_and_pre():
    g = torch.Generator(device="cuda").manual_seed(3)
    a = _mk(520, 768, False, g)
    w = (_mk(3072, 768, False, g).float() * 0.05).to(torch.bfloat16)
    bias = torch.randn(3072, generator=g, device="cuda") * 0.1
    pre = torch.empty((520, 3072), dtype=torch.bfloat16, device="cuda")
    out = ops.gemm(a, w, epilogue=ops.EPI_BIAS_GELU, bias=bias, c2=pre)
    ref_pre = a.float() @ w.float().t() + bias
    ref = torch.nn.functional.gelu(ref_pre)
    assert _rel(pre, ref_pre) < 1e-2   # bf16 output rounding
    assert _rel(out, ref) < 1e-2


def test_gemm_residual_dgelu_add():
    g = torch.Generator(device="cuda").manual_seed(4)
    a = _mk(260, 256, False, g)
    w = _mk(768, 256, False, g)
    bias = torch.randn(768, generator=g, device="cuda")
    res = _mk(260, 768, False, g)
    out = ops.gemm(a, w, epilogue=ops.EPI_BIAS_DROP_RES, bias=bias, aux=res)
    ref = a.float() @ w.float().t() + bias + res.float()
    assert _rel(out, ref) < 1e-2
    out = ops.gemm(a, w, epilogue=ops.EPI_ADD, aux=res)
    assert _rel(out, a.float() @ w.float().t() + res.float()) < 1e-2
    pre = (res.float() * 0.5).to(torch.bfloat16)
    out = ops.gemm(a, w, epilogue=ops.EPI_DGELU, aux=pre)
    x = pre.float().requires_grad_(True)
    torch.nn.functional.gelu(x).sum().backward()
    assert _rel(out, (a.float() @ w.float().t()) * x.grad) < 1e-2


def test_gemm_beta_accumulate():
    g = torch.Generator(device="cuda").manual_seed(5)
    a = _mk(128, 192, True, g)
    b = _mk(256, 192, True, g)
    c0 = torch.randn(128, 256, generator=g, device="cuda")
    out = c0.clone()
    ops.gemm(a, b, a_mn=True, b_mn=True, out=out, beta=1.0)
    assert _rel(out, a.float().t() @ b.float() + c0) < 2e-3


def test_gemm_dropout_consistent_and_unbiased():
    g = torch.Generator(device="cuda").manual_seed(6)
    a = _mk(512, 128, False, g)
    w = _mk(768, 128, False, g)
    res = torch.zeros((512, 768), dtype=torch.bfloat16, device="cuda")
    p = 0.1
    o1 = ops.gemm(a, w, epilogue=ops.EPI_BIAS_DROP_RES, aux=res, drop_p=p, seed=1234, site=2)
    o2 = ops.gemm(a, w, epilogue=ops.EPI_BIAS_DROP_RES, aux=res, drop_p=p, seed=1234, site=2)
    o3 = ops.gemm(a, w, epilogue=ops.EPI_BIAS_DROP_RES, aux=res, drop_p=p, seed=1235, site=2)
    assert torch.equal(o1, o2)
    assert not torch.equal(o1, o3)
    ref = a.float() @ w.float().t()
    kept = o1 != 0
    frac = 1.0 - kept.float().mean().item()
    assert abs(frac - p) < 0.01, frac
    assert _rel(o1.float()[kept], (ref / (1 - p))[kept]) < 1e-2


# ---- cta_group::2 pair kernel (block_n = 2000 + BN: a two-CTA cluster computes 256 x BN tiles) ----
@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K,bn", [(256, 256, 64, 2256), (256, 128, 64, 2128), (512, 512, 512, 2256),
                                       (1000, 768, 768, 2256), (1000, 768, 768, 2128), (304, 384, 192, 2256),
                                       (200, 136, 72, 2128), (4736, 1280, 320, 2256)])
def test_gemm_pair_majors(M, N, K, bn, a_mn, b_mn):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K + bn)
    a = _mk(M, K, a_mn, g)
    b = _mk(N, K, b_mn, g)
    out = ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, out_dtype=torch.float32, block_n=bn)
    ref = _ref(a, b, a_mn, b_mn)
    torch.cuda.synchronize()
    assert _rel(out, ref) < 2e-3, f"rel err {_rel(out, ref)}"


@pytest.mark.parametrize("bn", [2128, 2256, 0])
def test_gemm_pair_persistent_many_tiles_and_epilogues(bn):
    g = torch.Generator(device="cuda").manual_seed(11)
    a = _mk(9408, 768, False, g)
    w = (_mk(3072, 768, False, g).float() * 0.05).to(torch.bfloat16)
    bias = torch.randn(3072, generator=g, device="cuda") * 0.1
    pre = torch.empty((9408, 3072), dtype=torch.bfloat16, device="cuda")
    out = ops.gemm(a, w, epilogue=ops.EPI_BIAS_GELU, bias=bias, c2=pre, block_n=bn)
    ref_pre = a.float() @ w.float().t() + bias
    assert _rel(pre, ref_pre) < 1e-2
    assert _rel(out, torch.nn.functional.gelu(ref_pre)) < 1e-2
    # K = 3072 -> N = 768 with bias + residual, then the MN-major dgrad and the MN/MN wgrad of the same layer
    w2 = (_mk(768, 3072, False, g).float() * 0.05).to(torch.bfloat16)
    b2 = torch.randn(768, generator=g, device="cuda") * 0.1
    res = _mk(9408, 768, False, g)
    y = ops.gemm(out, w2, epilogue=ops.EPI_BIAS_DROP_RES, bias=b2, aux=res, block_n=bn)
    assert _rel(y, out.float() @ w2.float().t() + b2 + res.float()) < 1e-2
    dy = _mk(9408, 768, False, g)
    dh = ops.gemm(dy, w2, b_mn=True, block_n=bn)
    assert _rel(dh, dy.float() @ w2.float()) < 1e-2
    gw = ops.gemm(dy, out, a_mn=True, b_mn=True, out_dtype=torch.float32, block_n=bn)
    assert _rel(gw, dy.float().t() @ out.float()) < 2e-3


@pytest.mark.parametrize("bn", [2128, 2256])
def test_gemm_pair_splitk_and_beta(bn):
    g = torch.Generator(device="cuda").manual_seed(12)
    M, N, K = 768, 512, 4096
    a = _mk(M, K, True, g)
    b = _mk(N, K, True, g)
    ref = a.float().t() @ b.float()
    out = ops.gemm(a, b, a_mn=True, b_mn=True, out_dtype=torch.float32, splits=4, block_n=bn)
    assert _rel(out, ref) < 2e-3
    c0 = torch.randn(M, N, generator=g, device="cuda")
    acc = c0.clone()
    ops.gemm(a, b, a_mn=True, b_mn=True, out=acc, beta=1.0, block_n=bn)
    assert _rel(acc, ref + c0) < 2e-3


@pytest.mark.parametrize("bn", [128, 2256])
def test_gemm_dropout_mask_shared_by_forward_and_backward_epilogues(bn):
    """GELU+dropout (forward) and dGELU+dropout (dgrad) at the same (seed, site) drop the same elements, the mask does
    not depend on the tile shape, and the survivors are scaled by 1/(1-p)."""
    g = torch.Generator(device="cuda").manual_seed(13)
    M, N, K, p = 2304, 1024, 256, 0.1
    a = _mk(M, K, False, g)
    w = _mk(N, K, False, g)
    bias = torch.randn(N, generator=g, device="cuda")
    pre = torch.empty((M, N), dtype=torch.bfloat16, device="cuda")
    h = ops.gemm(a, w, epilogue=ops.EPI_BIAS_GELU, bias=bias, c2=pre, drop_p=p, seed=77, site=5, block_n=bn)
    h0 = ops.gemm(a, w, epilogue=ops.EPI_BIAS_GELU, bias=bias, block_n=bn)
    dy = _mk(M, K, False, g)
    dh = ops.gemm(dy, w, epilogue=ops.EPI_DGELU, aux=pre, drop_p=p, seed=77, site=5, block_n=bn)
    dh0 = ops.gemm(dy, w, epilogue=ops.EPI_DGELU, aux=pre, block_n=bn)
    live = (h0 != 0) & (dh0 != 0)
    keep_f = (h != 0)[live]
    keep_b = (dh != 0)[live]
    assert torch.equal(keep_f, keep_b)
    frac = 1.0 - keep_f.float().mean().item()
    assert abs(frac - p) < 0.005, frac
    assert _rel(h.float()[live & (h != 0)], (h0.float() / (1 - p))[live & (h != 0)]) < 1e-2
    # same mask from the single-CTA 128-wide kernel
    h_ref = ops.gemm(a, w, epilogue=ops.EPI_BIAS_GELU, bias=bias, drop_p=p, seed=77, site=5, block_n=128)
    assert torch.equal((h_ref != 0)[live], keep_f)


# ---- plain bf16 outputs of the pair kernel: TMA-store drain (default) / 256-bit stores (LR2_GEMM_DIRECT=1) ----
@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (True, True), (False, True)])
@pytest.mark.parametrize("M,N,K,pad", [(256, 256, 64, 0), (9408, 768, 192, 0), (304, 384, 192, 0), (1000, 768, 320, 64),
                                        (3072, 2560, 48, 0), (264, 264, 80, 24)])
def test_gemm_pair_plain_bf16_output_clipping_and_pitch(M, N, K, pad, a_mn, b_mn):
    """Ragged M and N (boxes clipped by the TMA unit), a row pitch larger than N (columns beside the output must stay
    untouched), K below one k-block (the out_layer.fc1 weight-gradient shape class)."""
    g = torch.Generator(device="cuda").manual_seed(M + 3 * N + 5 * K)
    a = _mk(M, K, a_mn, g)
    b = _mk(N, K, b_mn, g)
    full = torch.full((M + 3, N + pad), 7.0, device="cuda", dtype=torch.bfloat16)
    out = full[:M, :N]
    ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, out=out, block_n=2256)
    ref = _ref(a, b, a_mn, b_mn)
    torch.cuda.synchronize()
    assert _rel(out, ref) < 1e-2, f"rel err {_rel(out, ref)}"
    assert torch.equal(out, ref.to(torch.bfloat16)) or _rel(out, ref.to(torch.bfloat16).float()) < 8e-3
    assert (full[M:] == 7.0).all() and (full[:, N:] == 7.0).all()        # nothing written outside [M, N]
    # the staged single-CTA kernel produces the same bits (same fp32 accumulation order per k-block, same rounding)
    out1 = ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, block_n=128)
    assert _rel(out, out1.float()) < 8e-3
