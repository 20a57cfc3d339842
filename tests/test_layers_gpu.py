"""LayerNorm / XiT attention / glue kernels vs plain torch fp32 references of the same op (bf16 I/O:
tolerance 2e-2 relative to the tensor scale, north_star's bf16 bound)."""
import json
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from lr2ppo_b200 import ops
from oracle import restate

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ROWS = json.load(open(os.path.join(ROOT, "tests", "golden", "rows.json")))
bf = torch.bfloat16


def _rel(d, ref):
    return ((d.float() - ref.float()).abs().max() / ref.float().abs().max().clamp_min(1e-6)).item()


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("rows,D", [(9408, 768), (37, 768), (5, 64), (48, 1024)])
def test_layernorm_fwd_bwd(mode, rows, D):
    g = torch.Generator(device="cuda").manual_seed(rows + D + mode)
    x = (torch.randn(rows, D, generator=g, device="cuda") * 1.5 + 0.3).to(bf)
    gamma = 1 + 0.1 * torch.randn(D, generator=g, device="cuda")
    beta = 0.1 * torch.randn(D, generator=g, device="cuda")
    dy = torch.randn(rows, D, generator=g, device="cuda").to(bf)
    add = torch.randn(rows, D, generator=g, device="cuda").to(bf)
    eps = 1e-5 if mode == 0 else 1e-6
    xr = x.float().requires_grad_(True); gr = gamma.clone().requires_grad_(True); br = beta.clone().requires_grad_(True)
    yref = F.layer_norm(xr, (D,), gr, br, eps) if mode == 0 else restate.tencent_layernorm(xr, gr, br, eps)
    (yref * dy.float()).sum().backward()
    y, stats = ops.layernorm_fwd(x, gamma, beta, eps, mode)
    assert _rel(y, yref.detach()) < 1e-2
    dx, dxm, dgamma, dbeta = ops.layernorm_bwd(dy, x, gamma, stats, eps, mode, add=add, want_masked=True)
    assert _rel(dx, xr.grad + add.float()) < 1e-2
    assert torch.equal(dx, dxm)                                  # p = 0 -> masked copy is identical
    assert _rel(dgamma, gr.grad) < 5e-3 and _rel(dbeta, br.grad) < 5e-3


def test_layernorm_tencent_golden():
    c = ROWS["tencent_ln"]
    x = torch.tensor(c["x"], device="cuda").to(bf)
    gamma = torch.tensor(c["gamma"], device="cuda"); beta = torch.tensor(c["beta"], device="cuda")
    y, stats = ops.layernorm_fwd(x, gamma, beta, 1e-6, 1)
    assert _rel(y, torch.tensor(c["y"], device="cuda")) < 2e-2
    dx, _, dg, db = ops.layernorm_bwd(torch.tensor(c["gy"], device="cuda").to(bf), x, gamma, stats, 1e-6, 1)
    assert _rel(dx, torch.tensor(c["dx"], device="cuda")) < 2e-2
    assert _rel(dg, torch.tensor(c["dgamma"], device="cuda")) < 2e-2
    assert _rel(db, torch.tensor(c["dbeta"], device="cuda")) < 2e-2


def test_layernorm_regroup_and_dropout_mask_matches_gemm():
    g = torch.Generator(device="cuda").manual_seed(9)
    items, S, I, D = 3, 196, 16, 768
    x = torch.randn(items * S, D, generator=g, device="cuda").to(bf)
    gamma = torch.ones(D, device="cuda"); beta = torch.zeros(D, device="cuda")
    cat = torch.zeros(items * (S + I), D, dtype=bf, device="cuda")
    ops.layernorm_fwd(x, gamma, beta, 1e-5, 0, out=cat, regroup=(S, S + I, 0))
    y, stats = ops.layernorm_fwd(x, gamma, beta, 1e-5, 0)
    c3 = cat.view(items, S + I, D)
    assert torch.equal(c3[:, :S].reshape(-1, D), y) and (c3[:, S:] == 0).all()
    # masked gradient uses the same Philox stream as the GEMM epilogue that applied the dropout
    dy = torch.randn(items * S, D, generator=g, device="cuda").to(bf)
    p, seed, site = 0.1, 42, 3
    dx, dxm, _, _ = ops.layernorm_bwd(dy, x, gamma, stats, 1e-5, 0, drop_p=p, seed=seed, site=site, want_masked=True)
    eye = torch.eye(D, device="cuda").to(bf)
    ones = torch.ones(items * S, D, device="cuda").to(bf)
    zero = torch.zeros(items * S, D, dtype=bf, device="cuda")
    mask = ops.gemm(ones, eye, epilogue=ops.EPI_BIAS_DROP_RES, aux=zero, drop_p=p, seed=seed, site=site)  # keep/(1-p)
    assert _rel(dxm, dx.float() * mask.float()) < 1e-2
    assert ((dxm == 0) == (mask == 0)).float().mean() > 0.999


@pytest.mark.parametrize("items,Sq,Skv,H,E", [(5, 196, 16, 8, 768), (24, 2, 2, 8, 768), (24, 4, 4, 8, 768),
                                               (3, 50, 7, 4, 256),
                                               # tcgen05 path (heads of 96, <= 16 keys, 64..256 query rows)
                                               (48, 196, 16, 8, 768), (3, 64, 5, 8, 768), (2, 256, 16, 8, 768),
                                               (2, 130, 9, 8, 768), (1, 129, 1, 8, 768)])
def test_xattn_fwd_bwd(items, Sq, Skv, H, E):
    g = torch.Generator(device="cuda").manual_seed(items * Sq)
    q = (torch.randn(items, Sq, E, generator=g, device="cuda") * 0.3).to(bf)
    kv = (torch.randn(items, Skv, 2 * E, generator=g, device="cuda") * 0.3).to(bf)   # merged K|V buffer
    k, v = kv[:, :, :E], kv[:, :, E:]
    do = torch.randn(items, Sq, E, generator=g, device="cuda").to(bf)
    post = 1.0 / math.sqrt(E)
    qr, kr, vr = (t.float().requires_grad_(True) for t in (q, k, v))
    dh = E // H
    qh = qr.view(items, Sq, H, dh).transpose(1, 2); kh = kr.reshape(items, Skv, H, dh).transpose(1, 2)
    vh = vr.reshape(items, Skv, H, dh).transpose(1, 2)
    att = torch.softmax(qh @ kh.transpose(-1, -2), -1) * post          # finetune/xit.py:142-143
    oref = (att @ vh).transpose(1, 2).reshape(items, Sq, E)
    (oref * do.float()).sum().backward()
    o = ops.xattn_fwd(q, k, v, H, 1.0, post)
    assert _rel(o, oref.detach()) < 1e-2
    dkv = torch.empty_like(kv)
    dq, dk, dv = ops.xattn_bwd(q, k, v, do, H, 1.0, post, dkv_out=(dkv[:, :, :E], dkv[:, :, E:]))
    assert _rel(dq, qr.grad) < 2e-2 and _rel(dk, kr.grad) < 2e-2 and _rel(dv, vr.grad) < 2e-2


def test_glue_kernels():
    g = torch.Generator(device="cuda").manual_seed(2)
    bs, T, R = 6, 2, 196 * 768
    src = torch.randn(bs, T, R, generator=g, device="cuda")
    idx = torch.randint(0, T, (bs, 4), generator=g, device="cuda")
    out = ops.cast_gather(src, idx)
    ref = src[torch.arange(bs, device="cuda").view(bs, 1), idx].to(bf)
    assert torch.equal(out, ref)
    assert torch.equal(ops.cast_gather(src), src.to(bf))
    # concat placement + accumulate
    items, S, I, D = 4, 196, 16, 768
    img = torch.randn(items * I, D, generator=g, device="cuda").to(bf)
    cat = torch.zeros(items * (S + I), D, dtype=bf, device="cuda")
    ops.rows_copy(img, I, 0, cat, S + I, S, items, I, D)
    assert torch.equal(cat.view(items, S + I, D)[:, S:].reshape(-1, D), img)
    acc = img.clone()
    ops.rows_copy(cat, S + I, S, acc, I, 0, items, I, D, accumulate=True)
    assert _rel(acc, img.float() * 2) < 1e-2
    # column sums
    x = torch.randn(9408, 3072, generator=g, device="cuda").to(bf)
    assert _rel(ops.colsum(x), x.float().sum(0)) < 1e-4
    x = torch.randn(48, 768, generator=g, device="cuda").to(bf)
    assert _rel(ops.colsum(x), x.float().sum(0)) < 1e-4
    # head: dot with the last token of each group
    bsz, Tt = 24, 4
    xx = torch.randn(bsz * Tt, D, generator=g, device="cuda").to(bf)
    w = torch.randn(D, generator=g, device="cuda"); b = torch.randn(1, generator=g, device="cuda")
    out = ops.rowdot_fwd(xx, w, b, bsz, Tt, Tt - 1)
    ref = xx.float().view(bsz, Tt, D)[:, -1] @ w + b
    assert _rel(out, ref) < 1e-5
    dout = torch.randn(bsz, generator=g, device="cuda")
    dx, dw, db = ops.rowdot_bwd(xx, w, dout, bsz, Tt, Tt - 1)
    dref = torch.zeros(bsz, Tt, D, device="cuda"); dref[:, -1] = dout[:, None] * w[None]
    assert _rel(dx, dref.view(-1, D)) < 1e-2
    assert _rel(dw, (dout[:, None] * xx.float().view(bsz, Tt, D)[:, -1]).sum(0)) < 1e-5
    assert abs(db.item() - dout.sum().item()) < 1e-4
    # positional embedding
    pos = torch.randn(4, D, generator=g, device="cuda")
    y = xx.clone()
    ops.add_pos_fwd(y, pos, bsz, Tt)
    assert _rel(y.view(bsz, Tt, D), xx.float().view(bsz, Tt, D) + pos[None]) < 1e-2
    assert _rel(ops.add_pos_bwd(xx, bsz, Tt), xx.float().view(bsz, Tt, D).sum(0)) < 1e-5
    assert torch.equal(ops.to_f32(ops.to_bf16(src[0, 0])), src[0, 0].to(bf).float())
