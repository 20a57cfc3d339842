"""ProjectionLayer with the reference's constructor and state_dict keys (finetune/project_embedding.py:5-26):
Linear -> GELU -> Linear -> Dropout -> + projected -> LayerNorm, as three fused launches forward
(GEMM + bias + GELU with the pre-activation kept as the residual; GEMM + bias + dropout + residual; LayerNorm)."""
import torch
import torch.nn as nn

from . import engine as eng
from . import ops
from .ops import EPI_BIAS_GELU, EPI_BIAS_DROP_RES, EPI_DGELU, EPI_ADD

bf16 = torch.bfloat16


class _ProjFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, x, train, save, *params):
        shp = x.shape
        xb = x.detach().reshape(-1, shp[-1])
        xb = xb.contiguous() if xb.dtype == bf16 else ops.to_bf16(xb.float().contiguous())
        bank = mod._bank
        wp, wf = bank.get(mod.projection.weight), bank.get(mod.fc.weight)
        p = mod.dropout.p if train else 0.0
        seed = (torch.initial_seed() * 1000003 + id(mod) % 9973) & 0x7FFFFFFFFFFFFFFF
        sdev = None
        if p > 0.0:
            if mod._seed_counter is None or mod._seed_counter.device != x.device:
                mod._seed_counter = torch.zeros(1, dtype=torch.int64, device=x.device)
            ops.bump_counter(mod._seed_counter, 1)
            sdev = mod._seed_counter.clone() if save else mod._seed_counter
        projected = torch.empty((xb.shape[0], wp.shape[0]), dtype=bf16, device=x.device)
        g = ops.gemm(xb, wp, epilogue=EPI_BIAS_GELU, bias=mod.projection.bias.detach(), c2=projected)
        t = ops.gemm(g, wf, epilogue=EPI_BIAS_DROP_RES, bias=mod.fc.bias.detach(), aux=projected, drop_p=p, seed=seed,
                     site=1, seed_dev=sdev)
        ln = mod.layer_norm
        y, st = ops.layernorm_fwd(t, ln.weight.detach(), ln.bias.detach(), ln.eps, 0, want_stats=save)
        ctx.saved = (xb, projected, g, t, st, p, seed, sdev, wp, wf) if save else None
        ctx.mod, ctx.shp, ctx.dt, ctx.np, ctx.need_dx = mod, shp, x.dtype, len(params), x.requires_grad
        y = y.view(shp[:-1] + (y.shape[-1],))
        return y if x.dtype == bf16 else ops.to_f32(y)

    @staticmethod
    def backward(ctx, dy):
        mod = ctx.mod
        xb, projected, g, t, st, p, seed, sdev, wp, wf = ctx.saved
        d = dy.reshape(-1, dy.shape[-1])
        d = d.contiguous() if d.dtype == bf16 else ops.to_bf16(d.float().contiguous())
        sink = eng._GradSink()
        ln = mod.layer_norm
        dt, dtm, dg, db = ops.layernorm_bwd(d, t, ln.weight.detach(), st, ln.eps, 0, drop_p=p, seed=seed, site=1,
                                            want_masked=True, seed_dev=sdev)
        sink.put_vec(ln.weight, dg); sink.put_vec(ln.bias, db)
        eng._wgrad(sink, mod.fc, dtm, g)
        dgp = eng._dgrad(dtm, wf, epilogue=EPI_DGELU, aux=projected)       # through GELU'(projected)
        # d projected = dgp (GELU branch) + dt (residual branch)
        dproj = eng._add_bf16(dgp, dt)
        eng._wgrad(sink, mod.projection, dproj, xb)
        dx = None
        if ctx.need_dx:
            dx = eng._dgrad(dproj, wp).view(ctx.shp)
            dx = dx if ctx.dt == bf16 else ops.to_f32(dx)
        return (None, dx, None, None) + (None,) * ctx.np


class ProjectionLayer(nn.Module):
    def __init__(self, embedding_dim, projection_dim, dropout=0.2):
        super().__init__()
        self.projection = nn.Linear(embedding_dim, projection_dim)
        self.gelu = nn.GELU()
        self.fc = nn.Linear(projection_dim, projection_dim)
        self.dropout = nn.Dropout(dropout)
        self.layer_norm = nn.LayerNorm(projection_dim)
        self._bank = eng.ShadowBank()
        self._seed_counter = None

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("lr2ppo_b200.ProjectionLayer runs on CUDA (sm_100a) only; there is no CPU fallback")
        params = list(self.parameters())
        save = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in params))
        return _ProjFn.apply(self, x, self.training, save, *params)
