"""Fusion-model engine: explicit forward / backward of the LR2PPO fusion network as a fixed sequence of
C-ABI kernel launches on bf16 activations (fp32 master weights, bf16 shadow weights, fp32 gradients).

It runs the computation of the reference's Actor / Critic / Reward / Classifier forward
(finetune/ppo.py:214-244, 265-297, 318-350; finetune/pointwise.py:207-236;
finetune/reward_pair_dataloader.py:251-283) and of torch autograd's backward through them:

    text_proj, img_proj (Mlp)      -> tcgen05 GEMM + bias + erf-GELU epilogues
    XiT block                      -> LayerNorm kernels, Q/K/V/O GEMMs, tiny-KV attention kernel,
                                      FFN GEMMs with GELU / dropout / residual epilogues
    cat + out_layer (162816->3072) -> the final LayerNorm writes straight into the concat buffer;
                                      fc1 is a swapped, split-K, transposed-output GEMM that streams
                                      the 1 GB bf16 weight once
    pos_emb + xitt + head          -> same kernels on [bs, T<=4] tokens

Nothing here computes on the host and there is no fallback path.
"""
import math

import os

import torch

from . import ops
from .ops import EPI_NONE, EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_DROP_RES, EPI_DGELU, EPI_ADD

bf16 = torch.bfloat16
f32 = torch.float32

XIT_HEADS = 8          # finetune/xit.py:114 (num_heads default)
XIT_DROP = 0.1         # finetune/xit.py:26,28 (drop_p, forward_drop_p)
LN_EPS = 1e-5          # nn.LayerNorm default


class ShadowBank:
    """bf16 copies of fp32 parameters, re-cast when the parameter was modified (tracked with Tensor._version).
    FusedAdamW refreshes the shadows that are registered with it in its own pass and leaves their parameter's
    version alone; for every other parameter it updates through raw pointers it bumps the version, so the copy
    held here is re-cast at the next `get`.  A re-cast always lands in the SAME bf16 tensor (shape permitting), so
    tensors registered with an optimizer / captured in a CUDA graph stay the ones forward reads."""

    def __init__(self):
        self._sh = {}

    def invalidate(self):
        """Mark every copy stale (e.g. after `p.data` was overwritten by a broadcast or a checkpoint load, which do
        not bump the version counter): the next `get` re-casts in place into the existing bf16 tensor."""
        self._sh = {k: (sh, None, ptr) for k, (sh, _, ptr) in self._sh.items()}

    def get(self, p):
        ent = self._sh.get(id(p))
        if ent is None or ent[1] != p._version or ent[2] != p.data_ptr():
            sh = ent[0] if ent is not None and ent[0].shape == p.shape and ent[0].device == p.device else \
                torch.empty(p.shape, dtype=bf16, device=p.device)
            ops.to_bf16(p.detach().contiguous(), out=sh)
            ent = (sh, p._version, p.data_ptr())
            self._sh[id(p)] = ent
        return ent[0]

    def items(self):
        return self._sh.items()


class _Lin:
    """weight [out, in] (bf16 shadow) + fp32 bias of one nn.Linear."""

    def __init__(self, bank, lin):
        self.mod = lin
        self.w = bank.get(lin.weight)
        self.b = lin.bias.detach()


class _GradSink:
    """Where backward puts parameter gradients.

    written=None: torch semantics — allocate when `.grad` is None, otherwise accumulate.
    written=set : persistent-buffer mode (CUDA-graph friendly): `.grad` tensors are kept across steps; the first
                  write to a parameter since `FusionEngine.begin_step()` overwrites the stale value, later writes
                  in the same step (second forward/backward pair of stage 2) accumulate."""

    def __init__(self, written=None):
        self.written = written

    def _stale(self, p):
        if self.written is None or id(p) in self.written:
            return False
        self.written.add(id(p))
        return True

    def buf(self, p):
        if p.grad is None:
            p.grad = torch.empty_like(p, memory_format=torch.contiguous_format)
            if self.written is not None:
                self.written.add(id(p))
            return p.grad, 0.0
        return p.grad, (0.0 if self._stale(p) else 1.0)

    def put_vec(self, p, val):
        """val: fp32 tensor shaped like p (small vectors: biases, LayerNorm params, head)."""
        if p.grad is None:
            p.grad = val.reshape(p.shape).clone()
            if self.written is not None:
                self.written.add(id(p))
        elif self._stale(p):
            p.grad.copy_(val.reshape(p.shape))
        else:
            p.grad.add_(val.reshape(p.shape))


def _wgrad(sink, lin, dy, x, bias_from=None):
    """dW[out,in] (+)= dy[M,out]^T @ x[M,in]; db (+)= colsum(dy).  Both operands MN-major -> no transposes."""
    gw, beta = sink.buf(lin.weight)
    m_out, k_in = lin.weight.shape
    rows = dy.shape[0]
    # few output tiles and a long reduction -> split-K to fill the SMs (CTA pairs when the pair kernel applies)
    splits = ops.plan_gemm(m_out, k_in, rows) if beta == 0.0 else 1
    tiles = ((m_out + 127) // 128) * ((k_in + 127) // 128)
    if splits == 1 and beta == 0.0 and tiles < 120 and rows >= 1024 and \
            not (k_in % 256 == 0 and m_out >= 256 and ((m_out + 255) // 256) * (k_in // 256) >= 37):
        splits = max(1, min(16, 148 // tiles, rows // 512))
    ops.gemm(dy, x, a_mn=True, b_mn=True, out=gw, beta=beta, splits=splits)
    sink.put_vec(lin.bias, ops.colsum(dy if bias_from is None else bias_from))


def _dgrad(dy, lin_w, **kw):
    """dx[M,in] = dy[M,out] @ W[out,in]   (W is the MN-major B operand)."""
    return ops.gemm(dy, lin_w, b_mn=True, **kw)


class XitWeights:
    def __init__(self, bank, xit):
        blk = xit[0][0]
        att_seq, ffn_seq = blk[0].fn, blk[1].fn
        self.lnb = att_seq[0]
        mha = att_seq[1]
        self.mha = mha
        self.heads = mha.num_heads
        self.emb = mha.emb_size
        self.q, self.k, self.v, self.o = (_Lin(bank, m) for m in (mha.queries, mha.keys, mha.values, mha.projection))
        self.ln2 = ffn_seq[0]
        self.f1, self.f2 = _Lin(bank, ffn_seq[1][0]), _Lin(bank, ffn_seq[1][3])
        self.ln3 = xit[1][0]
        self.p_att = att_seq[2].p
        self.p_ffn_in = ffn_seq[1][2].p
        self.p_ffn_out = ffn_seq[2].p


def xit_forward(W, x, y, items, Sq, Skv, train, seed, site_base, save, out=None, regroup=None, seed_dev=None):
    """XiT block on x [items*Sq, E] (queries) and y [items*Skv, E] (keys/values).
    ref: finetune/xit.py:9-148.  Returns (output or `out`, ctx)."""
    E = W.emb
    p1 = W.p_att if train else 0.0
    p2 = W.p_ffn_in if train else 0.0
    p3 = W.p_ffn_out if train else 0.0
    lx, st_x = ops.layernorm_fwd(x, W.lnb.ln_x.weight.detach(), W.lnb.ln_x.bias.detach(), W.lnb.ln_x.eps, 0,
                                 want_stats=save)
    ly, st_y = ops.layernorm_fwd(y, W.lnb.ln_y.weight.detach(), W.lnb.ln_y.bias.detach(), W.lnb.ln_y.eps, 0,
                                 want_stats=save)
    q = ops.gemm(lx, W.q.w, epilogue=EPI_BIAS, bias=W.q.b)
    k = ops.gemm(ly, W.k.w, epilogue=EPI_BIAS, bias=W.k.b)
    v = ops.gemm(ly, W.v.w, epilogue=EPI_BIAS, bias=W.v.b)
    post = 1.0 / math.sqrt(E)     # softmax first, then / sqrt(emb): finetune/xit.py:142-143
    a = ops.xattn_fwd(q.view(items, Sq, E), k.view(items, Skv, E), v.view(items, Skv, E), W.heads, 1.0, post)
    a = a.view(items * Sq, E)
    x1 = ops.gemm(a, W.o.w, epilogue=EPI_BIAS_DROP_RES, bias=W.o.b, aux=x, drop_p=p1, seed=seed, site=site_base + 1,
                  seed_dev=seed_dev)
    l2, st2 = ops.layernorm_fwd(x1, W.ln2.weight.detach(), W.ln2.bias.detach(), W.ln2.eps, 0, want_stats=save)
    pre2 = torch.empty((items * Sq, W.f1.w.shape[0]), dtype=bf16, device=x.device) if save else None
    h2 = ops.gemm(l2, W.f1.w, epilogue=EPI_BIAS_GELU, bias=W.f1.b, c2=pre2, drop_p=p2, seed=seed, site=site_base + 2,
                  seed_dev=seed_dev)
    x2 = ops.gemm(h2, W.f2.w, epilogue=EPI_BIAS_DROP_RES, bias=W.f2.b, aux=x1, drop_p=p3, seed=seed,
                  site=site_base + 3, seed_dev=seed_dev)
    xo, st3 = ops.layernorm_fwd(x2, W.ln3.weight.detach(), W.ln3.bias.detach(), W.ln3.eps, 0, out=out,
                                regroup=regroup, want_stats=save)
    ctx = None
    if save:
        ctx = dict(x=x, y=y, lx=lx, ly=ly, st_x=st_x, st_y=st_y, q=q, k=k, v=v, a=a, x1=x1, st2=st2, l2=l2,
                   pre2=pre2, h2=h2, x2=x2, st3=st3, p=(p1, p2, p3), seed=seed, seed_dev=seed_dev, site_base=site_base,
                   dims=(items, Sq, Skv), regroup=regroup)
    return xo, ctx


def xit_backward(W, ctx, dout, sink, need_dx=True, need_dy=True, dy_extra=None):
    """Backward of xit_forward.  dout: gradient wrt the block output (row-regrouped like the output).
    Returns (dx, dy): gradients wrt the query-side and key/value-side inputs (dy includes dy_extra)."""
    items, Sq, Skv = ctx["dims"]
    E = W.emb
    p1, p2, p3 = ctx["p"]
    seed, sb, sdev = ctx["seed"], ctx["site_base"], ctx["seed_dev"]
    post = 1.0 / math.sqrt(E)
    # final LayerNorm; dx2m = gradient into the (dropout-ed) FFN output
    dx2, dx2m, dg, db = ops.layernorm_bwd(dout, ctx["x2"], W.ln3.weight.detach(), ctx["st3"], W.ln3.eps, 0,
                                          regroup=ctx["regroup"], drop_p=p3, seed=seed, site=sb + 3,
                                          want_masked=True, seed_dev=sdev)
    sink.put_vec(W.ln3.weight, dg); sink.put_vec(W.ln3.bias, db)
    # FFN
    _wgrad(sink, W.f2.mod, dx2m, ctx["h2"])
    dh2p = _dgrad(dx2m, W.f2.w, epilogue=EPI_DGELU, aux=ctx["pre2"], drop_p=p2, seed=seed, site=sb + 2,
                  seed_dev=sdev)
    _wgrad(sink, W.f1.mod, dh2p, ctx["l2"])
    dl2 = _dgrad(dh2p, W.f1.w)
    dx1, dx1m, dg, db = ops.layernorm_bwd(dl2, ctx["x1"], W.ln2.weight.detach(), ctx["st2"], W.ln2.eps, 0, add=dx2,
                                          drop_p=p1, seed=seed, site=sb + 1, want_masked=True,
                                          seed_dev=sdev)
    sink.put_vec(W.ln2.weight, dg); sink.put_vec(W.ln2.bias, db)
    # attention
    _wgrad(sink, W.o.mod, dx1m, ctx["a"])
    da = _dgrad(dx1m, W.o.w)
    dq, dk, dv = ops.xattn_bwd(ctx["q"].view(items, Sq, E), ctx["k"].view(items, Skv, E),
                               ctx["v"].view(items, Skv, E), da.view(items, Sq, E), W.heads, 1.0, post)
    dq = dq.view(items * Sq, E); dk = dk.view(items * Skv, E); dv = dv.view(items * Skv, E)
    _wgrad(sink, W.q.mod, dq, ctx["lx"])
    _wgrad(sink, W.k.mod, dk, ctx["ly"])
    _wgrad(sink, W.v.mod, dv, ctx["ly"])
    dxin = dyin = None
    if need_dy:
        dly = _dgrad(dk, W.k.w)
        dly = _dgrad(dv, W.v.w, epilogue=EPI_ADD, aux=dly)
        dyin, _, dg, db = ops.layernorm_bwd(dly, ctx["y"], W.lnb.ln_y.weight.detach(), ctx["st_y"], W.lnb.ln_y.eps, 0,
                                            add=dy_extra)
    else:
        dly = _dgrad(dk, W.k.w)
        dly = _dgrad(dv, W.v.w, epilogue=EPI_ADD, aux=dly)
        _, _, dg, db = ops.layernorm_bwd(dly, ctx["y"], W.lnb.ln_y.weight.detach(), ctx["st_y"], W.lnb.ln_y.eps, 0)
    sink.put_vec(W.lnb.ln_y.weight, dg); sink.put_vec(W.lnb.ln_y.bias, db)
    dlx = _dgrad(dq, W.q.w)
    dxin, _, dg, db = ops.layernorm_bwd(dlx, ctx["x"], W.lnb.ln_x.weight.detach(), ctx["st_x"], W.lnb.ln_x.eps, 0,
                                        add=dx1)
    sink.put_vec(W.lnb.ln_x.weight, dg); sink.put_vec(W.lnb.ln_x.bias, db)
    return dxin, dyin


def mlp_forward(l1, l2, x, save):
    """Mlp: fc2(gelu(fc1 x)).  ref: finetune/ppo.py:164-170 (drop p = 0)."""
    pre = torch.empty((x.shape[0], l1.w.shape[0]), dtype=bf16, device=x.device) if save else None
    h = ops.gemm(x, l1.w, epilogue=EPI_BIAS_GELU, bias=l1.b, c2=pre)
    y = ops.gemm(h, l2.w, epilogue=EPI_BIAS, bias=l2.b)
    return y, (x, pre, h) if save else None


def mlp_backward(l1, l2, ctx, dy, sink, need_dx=False):
    x, pre, h = ctx
    _wgrad(sink, l2.mod, dy, h)
    dhp = _dgrad(dy, l2.w, epilogue=EPI_DGELU, aux=pre)
    _wgrad(sink, l1.mod, dhp, x)
    return _dgrad(dhp, l1.w) if need_dx else None


class FusionEngine:
    """Forward / backward of one fusion model (`kind` in actor | critic): `module` supplies the fp32
    parameters under the reference's attribute names (text_proj, img_proj, xit, out_layer, head
    [, pos_emb, xitt])."""

    def __init__(self, module, kind):
        self.m = module
        self.kind = kind
        self.bank = ShadowBank()
        self._calls = 0
        self.seed_counter = None   # int64[1] device tensor: dropout seed offset
        self.dropout_seed = None   # int: dropout seed of the NEXT training forward, incremented by one after each
                                   # (mask replay / parity tests: the masks become a known function of this value,
                                   # oracle/philox.py); None = torch.initial_seed()-derived base + the device counter
        self.persistent_grads = False  # keep .grad buffers across steps (see _GradSink); call begin_step() per step
        self._written = set()
        self.fc1_stash = None      # list of (dY, X) when out_layer.fc1 is updated by the fused wgrad+AdamW kernel
        self.dp_gather = None      # optional callable(t) -> all-gathered rows (data-parallel fused mode)
        self.dp_gather_async = None  # optional callable(t) -> handle; handle() waits and returns the gathered rows
        self.fc1_rows = None         # (r0, r1): rows of out_layer.fc1 this rank owns (row-sharded optimizer)
        self.fc1_grad_bf16 = None  # bf16 [out, in] gradient buffer of out_layer.fc1.weight (see enable_bf16_fc1_grad)
        self.fc1_passes, self._fc1_pending = 1, []   # backward passes per optimizer step (stage 2: 2) and their operands
        self._zero_index = {}      # (bs, T, device) -> int64 zeros [bs, T]: broadcast index of un-repeated img_emb
        self.tp = None             # dist.Fc1Parallel: out_layer.fc1 K-split over the data-parallel ranks
        self.on_trunk_grads = None  # dist.GradSync: callback fired inside backward once all but the projections' grads are final
        self._bwd_calls = 0        # backward passes since begin_step()

    def begin_step(self):
        """Persistent-gradient mode: start of a new optimizer step (replaces model.zero_grad())."""
        self._written.clear()
        self._fc1_pending = []
        self._bwd_calls = 0

    def _fc1_wgrad(self, dy_, x_, out, bn=None):
        """out (bf16 block of the out_layer.fc1 gradient) = dY^T X over all rows collected for this optimizer step.
        With fc1_passes == 2 (stage 2: chosen and reject forwards, one backward pass each) the operands of the first
        pass are stashed and the second pass runs ONE GEMM over both -- K = 2 x rows -- instead of writing a 2 GB fp32
        gradient and read-modify-writing it again."""
        if self.fc1_passes > 1:
            self._fc1_pending.append((dy_, x_))
            if len(self._fc1_pending) < self.fc1_passes:
                return
            dy_ = torch.cat([a for a, _ in self._fc1_pending], dim=0)
            x_ = torch.cat([b for _, b in self._fc1_pending], dim=0)
            self._fc1_pending = []
        if bn is None:
            # K = rows: a pure output-write problem (1 GB of bf16 per 500 M-parameter matrix).  Pair kernel with the
            # TMA-store drain: 215 us against 261 us for single-CTA 128-wide tiles with the staged drain (K = 48,
            # profiles/r02_gemm_shapes.txt); LR2_FC1_WGRAD_BN overrides (128 / 256 = the round-1 choices)
            bn = int(os.environ.get("LR2_FC1_WGRAD_BN", "2256"))
        ops.gemm(dy_, x_, a_mn=True, b_mn=True, out=out, block_n=bn)

    def enable_bf16_fc1_grad(self, optimizer, passes=1):
        """Keep the gradient of out_layer.fc1.weight (97 % of the parameters) in a persistent bf16 buffer that
        FusedAdamW reads directly; `.grad` of that parameter stays None.  One backward per optimizer step only
        (no accumulation), i.e. stage 1 and stage 3."""
        w = self.m.out_layer.fc1.weight
        self.fc1_passes, self._fc1_pending = int(passes), []
        self.fc1_grad_bf16 = torch.empty(w.shape, dtype=bf16, device=w.device)
        optimizer.register_shadow(w, self.bank.get(w))
        optimizer.register_grad(w, self.fc1_grad_bf16)

    def enable_fused_fc1(self, optimizer):
        """Route out_layer.fc1.weight through lr2_gemm_wgrad_adamw (gradient never materialised)."""
        self.fc1_stash = []
        w = self.m.out_layer.fc1.weight

        def provider():
            pairs, self.fc1_stash = self.fc1_stash, []
            if self.dp_gather is not None:
                pairs = [(self.dp_gather(a), self.dp_gather(b)) for a, b in pairs]
            return pairs

        optimizer.register_shadow(w, self.bank.get(w))
        optimizer.register_fused_wgrad(w, provider)

    # weights are re-wrapped per call (cheap) so that parameter updates are always seen
    def _weights(self):
        m, bank = self.m, self.bank
        W = dict(tp1=_Lin(bank, m.text_proj.fc1), tp2=_Lin(bank, m.text_proj.fc2),
                 ip1=_Lin(bank, m.img_proj.fc1), ip2=_Lin(bank, m.img_proj.fc2),
                 xit=XitWeights(bank, m.xit), o1=_Lin(bank, m.out_layer.fc1), o2=_Lin(bank, m.out_layer.fc2))
        if self.kind == "critic":
            W["xitt"] = XitWeights(bank, m.xitt)
        return W

    def forward(self, text, img, index=None, train=False, save=False, seed=None, trunk_only=False):
        """text [bs, Tsrc, S, E] fp32, img [bs, Tsrc, I, E] fp32, index [bs, T] int64 or None.
        Returns (logits fp32, ctx).  actor: logits [bs*T, n_out]; critic: [bs].
        trunk_only (inference, index None): stop after out_layer and return (pooled item features [bs*Tsrc, E] bf16,
        None) -- everything of a critic / reward forward that does not depend on the index; tail() finishes it."""
        m = self.m
        W = self._weights()
        bs, Tsrc, S, E = text.shape
        I = img.shape[2]
        T = index.shape[1] if index is not None else Tsrc
        # Inference-only reuse: when `index` repeats items (reward model: [0, 1, pi(0), pi(1)] holds 2 distinct
        # items) the pooled feature of each (clip, tag) item is computed once and gathered afterwards. Exact in
        # eval mode; not used when training (independent dropout masks per occurrence).
        reuse = index is not None and T > Tsrc and not train and not save
        body_index = None if reuse else index
        items = bs * (Tsrc if reuse else T)
        seed_dev = None
        if seed is None and self.dropout_seed is not None:
            seed = int(self.dropout_seed)
            if train:
                self.dropout_seed = seed + 1
        if seed is None:
            # dropout seed = host base (torch.initial_seed) + a device-resident counter bumped per training
            # forward, so a CUDA-graph replay of this call sequence still draws fresh masks; backward reads the
            # snapshot taken here.
            seed = (torch.initial_seed() * 1000003 + id(self) % 9973) & 0x7FFFFFFFFFFFFFFF
            if train:
                if self.seed_counter is None or self.seed_counter.device != text.device:
                    self.seed_counter = torch.zeros(1, dtype=torch.int64, device=text.device)
                ops.bump_counter(self.seed_counter, 1)
                seed_dev = self.seed_counter.clone() if save else self.seed_counter
        xt = _to_items(text.reshape(bs, Tsrc, S * E), body_index).view(items * S, E)
        if img.shape[1] == 1 and Tsrc > 1:
            # img_emb as the loader yields it, [bs, 1, I, E]: one keyframe set per clip shared by all of its tags.
            # The reference materialises img_emb.unsqueeze(1).repeat(1, tags, 1, 1) on the host and uploads the copies
            # (finetune/ppo.py:831); here the broadcast happens inside the gather + cast kernel (all-zero tag index).
            T_items = items // bs
            key = (bs, T_items, str(img.device))
            zi = self._zero_index.get(key)
            if zi is None:
                zi = self._zero_index[key] = torch.zeros((bs, T_items), dtype=torch.int64, device=img.device)
            xi = _to_items(img.reshape(bs, 1, I * E), zi).view(items * I, E)
        else:
            xi = _to_items(img.reshape(bs, Tsrc, I * E), body_index).view(items * I, E)
        tf, c_tp = mlp_forward(W["tp1"], W["tp2"], xt, save)
        imf, c_ip = mlp_forward(W["ip1"], W["ip2"], xi, save)
        cat = torch.empty((items, (S + I) * E), dtype=bf16, device=text.device)
        cat_rows = cat.view(items * (S + I), E)
        _, c_x = xit_forward(W["xit"], tf, imf, items, S, I, train, seed, 0, save, out=cat_rows,
                             regroup=(S, S + I, 0), seed_dev=seed_dev)
        ops.rows_copy(imf, I, 0, cat_rows, S + I, S, items, I, E)
        cat_all = None
        o1 = W["o1"]
        hid = o1.w.shape[0]
        tp = self.tp if (self.tp is not None and self.tp.active) else None
        if tp is not None:
            # out_layer.fc1 split along its input dimension over the data-parallel ranks (dist.Fc1Parallel): this rank
            # holds (and updates) only the column block W[:, k0:k1), multiplies that block of EVERY rank's rows and
            # the partial sums are reduce-scattered -- ~20 MB exchanged per forward instead of all-gathering the
            # updated 1 GB weight after every optimizer step.  HBM: the weight stream shrinks by `world` too.
            y1, pre3, cat_all = self._fc1_forward_tp(tp, o1, cat, items, save)
        else:
            if save and self.dp_gather_async is not None and self.fc1_grad_bf16 is not None:
                cat_all = self.dp_gather_async(cat)      # wgrad operand X of out_layer.fc1, consumed in backward
            # out_layer.fc1: weight is the 128-row MMA operand, items are N; split-K streams the weight once
            pre3 = torch.empty((items, hid), dtype=bf16, device=text.device) if save else None
            if items <= 256:
                bn = 64 if items <= 64 else (128 if items <= 128 else 256)
                kblocks = (cat.shape[1] + 63) // 64
                m_tiles = (hid + 127) // 128
                splits = max(1, min(kblocks, 148 // m_tiles))
                y1 = torch.empty((items, hid), dtype=bf16, device=text.device)
                ops.gemm(o1.w, cat, out=y1, transposed_out=True, epilogue=EPI_BIAS_GELU, bias=o1.b, c2=pre3,
                         splits=splits, block_n=bn)
            else:
                y1 = ops.gemm(cat, o1.w, epilogue=EPI_BIAS_GELU, bias=o1.b, c2=pre3, splits=4)
        feat = ops.gemm(y1, W["o2"].w, epilogue=EPI_BIAS, bias=W["o2"].b)
        if trunk_only:
            if train or save or index is not None:
                raise RuntimeError("trunk_only is the inference-only first half of forward(index=...)")
            return feat, None
        if reuse:
            feat = ops.gather_rows(feat.view(bs, Tsrc, E), index).view(bs * T, E)
            items = bs * T
        ctx = None
        if self.kind == "actor":
            n_out = m.head.weight.shape[0]
            if n_out == 1:
                logits = ops.rowdot_fwd(feat, m.head.weight.detach().view(-1), m.head.bias.detach(), items)
            else:
                logits = _small_linear(feat, m.head, self.bank)
            if save:
                ctx = dict(W=W, dims=(bs, T, S, I, E, items), c_tp=c_tp, c_ip=c_ip, c_x=c_x, cat=cat, pre3=pre3,
                           y1=y1, feat=feat, imf=imf, cat_all=cat_all, tp=tp)
            return logits, ctx
        # critic / reward: + pos_emb, self-attention over the T items, head on the LAST token
        ops.add_pos_fwd(feat, m.pos_emb.weight.detach()[:T].contiguous(), bs, T)
        z, c_t = xit_forward(W["xitt"], feat, feat, bs, T, T, train, seed, 3, save, seed_dev=seed_dev)
        logits = ops.rowdot_fwd(z, m.head.weight.detach().view(-1), m.head.bias.detach(), bs, T, T - 1)
        if save:
            ctx = dict(W=W, dims=(bs, T, S, I, E, items), c_tp=c_tp, c_ip=c_ip, c_x=c_x, cat=cat, pre3=pre3, y1=y1,
                       feat=feat, imf=imf, c_t=c_t, z=z, cat_all=cat_all, tp=tp)
        return logits, ctx

    def tail(self, feat_items, index):
        """Second half of an inference forward of a critic / reward model from the pooled item features of
        forward(trunk_only=True): gather by index [bs, T], + pos_emb, self-attention over the T items, head on the last
        token -- exactly the operations forward(index=...) runs after out_layer in its item-reuse mode."""
        m = self.m
        W = self._weights()
        bs, T = index.shape
        E = feat_items.shape[-1]
        Tsrc = feat_items.shape[0] // bs
        feat = ops.gather_rows(feat_items.view(bs, Tsrc, E), index).view(bs * T, E)
        ops.add_pos_fwd(feat, m.pos_emb.weight.detach()[:T].contiguous(), bs, T)
        z, _ = xit_forward(W["xitt"], feat, feat, bs, T, T, False, 0, 3, False)
        return ops.rowdot_fwd(z, m.head.weight.detach().view(-1), m.head.bias.detach(), bs, T, T - 1)

    def _fc1_forward_tp(self, tp, o1, cat, items, save):
        """K-split out_layer.fc1 forward (dist.Fc1Parallel).  Returns (y1 [items, hid], pre3 | None, X block | None)."""
        hid = o1.w.shape[0]
        k0, k1 = tp.cols
        total = tp.world * items
        x_k = tp.scatter_cols(cat)                                   # [world * items, Kb]: my column block, all rows
        w_k = o1.w[:, k0:k1]                                         # strided view, pitch K1
        part = torch.empty((total, hid), dtype=f32, device=cat.device)
        bn = 64 if total <= 64 else (128 if (total <= 128 or total > 256) else 256)
        tiles = ((hid + 127) // 128) * ((total + bn - 1) // bn)
        splits = max(1, min(((k1 - k0) + 63) // 64, 148 // tiles))
        ops.gemm(w_k, x_k, out=part, transposed_out=True, splits=splits, block_n=bn)
        y_sum = tp.reduce_scatter(part)                              # [items, hid] fp32: sum over the column blocks
        y1, pre3 = ops.bias_gelu_rows(y_sum, o1.b, want_pre=save)
        return y1, pre3, (x_k if save else None)

    def _fc1_backward_tp(self, tp, o1, dy1p, x_k, items):
        """K-split out_layer.fc1 backward: dX of the own column block for everybody's rows (complete sums over the
        hidden units), returned to the owners by an all-to-all that overlaps the weight-gradient GEMM of the block."""
        k0, k1 = tp.cols
        w_k = o1.w[:, k0:k1]
        dy_all = tp.all_gather(dy1p)                                 # [world * items, hid] (0.3 MB per rank)
        dx_k = ops.gemm(dy_all, w_k, b_mn=True)                      # [world * items, Kb]
        wait_dx = tp.gather_cols_async(dx_k)
        self._fc1_wgrad(dy_all, x_k, self.fc1_grad_bf16[:, k0:k1])
        return wait_dx().contiguous()                                # [items, K1]

    def backward(self, ctx, dlogits):
        """Accumulates fp32 gradients into module.parameters().grad (allocating when None)."""
        m = self.m
        W = ctx["W"]
        bs, T, S, I, E, items = ctx["dims"]
        sink = _GradSink(self._written if self.persistent_grads else None)
        dlogits = dlogits.contiguous().to(f32)
        if self.kind == "actor":
            if m.head.weight.shape[0] == 1:
                dfeat, dw, dbias = ops.rowdot_bwd(ctx["feat"], m.head.weight.detach().view(-1), dlogits.view(-1),
                                                  items)
                sink.put_vec(m.head.weight, dw); sink.put_vec(m.head.bias, dbias)
            else:
                dfeat = _small_linear_bwd(ctx["feat"], m.head, self.bank, dlogits.view(items, -1), sink)
        else:
            dz, dw, dbias = ops.rowdot_bwd(ctx["z"], m.head.weight.detach().view(-1), dlogits.view(-1), bs, T, T - 1)
            sink.put_vec(m.head.weight, dw); sink.put_vec(m.head.bias, dbias)
            dxa, dya = xit_backward(W["xitt"], ctx["c_t"], dz, sink)
            # x == y for xitt: total input gradient = dxa + dya
            dfeat = _add_bf16(dxa, dya)
            dpos = ops.add_pos_bwd(dfeat, bs, T)
            gp = torch.zeros_like(m.pos_emb.weight)
            gp[:T] = dpos
            sink.put_vec(m.pos_emb.weight, gp)
        # out_layer
        _wgrad(sink, W["o2"].mod, dfeat, ctx["y1"])
        dy1p = _dgrad(dfeat, W["o2"].w, epilogue=EPI_DGELU, aux=ctx["pre3"])
        deferred_fc1 = None
        tp = ctx.get("tp")
        if tp is not None:
            sink.put_vec(W["o1"].mod.bias, ops.colsum(dy1p))      # own items only: summed over ranks with the small grads
            dcat = self._fc1_backward_tp(tp, W["o1"], dy1p, ctx["cat_all"], items)
        else:
            dcat = None
        if tp is not None:
            pass
        elif self.fc1_stash is not None:
            # fused mode: the optimizer consumes (dY, X) in lr2_gemm_wgrad_adamw; no 2 GB gradient is written
            self.fc1_stash.append((dy1p, ctx["cat"]))
            sink.put_vec(W["o1"].mod.bias, ops.colsum(dy1p))
        elif self.fc1_grad_bf16 is not None:
            # bf16 gradient side-buffer read directly by FusedAdamW (1 GB written + read instead of 2 GB); in
            # data-parallel runs the two small wgrad operands are all-gathered (global-batch gradient, no all-reduce).
            # The wgrad GEMM itself is deferred to the END of this backward (nothing downstream reads the weight
            # gradient): the asynchronous gathers of dY (started here) and of X (started in forward) then have the
            # whole backward to complete instead of sitting in front of it.
            fc1_dy_handle = self.dp_gather_async(dy1p) if self.dp_gather_async is not None else None
            deferred_fc1 = (dy1p, fc1_dy_handle)
            sink.put_vec(W["o1"].mod.bias, ops.colsum(dy1p))
        elif self.dp_gather is not None:
            # data parallel: gather the two (small) wgrad operands instead of all-reducing the 2 GB gradient
            _wgrad(sink, W["o1"].mod, self.dp_gather(dy1p), self.dp_gather(ctx["cat"]), bias_from=dy1p)
        else:
            _wgrad(sink, W["o1"].mod, dy1p, ctx["cat"])
        if dcat is not None:
            pass
        elif items <= 256:
            dcat = torch.empty_like(ctx["cat"])
            bn = 64 if items <= 64 else (128 if items <= 128 else 256)
            # dcat^T[K1, items] = W1^T[K1, hid] @ dy1p^T : W1 is the MN-major A operand, output written transposed
            ops.gemm(W["o1"].w, dy1p, a_mn=True, out=dcat, transposed_out=True, block_n=bn)
        else:
            dcat = torch.empty_like(ctx["cat"])
            ops.gemm(dy1p, W["o1"].w, b_mn=True, out=dcat)
        dcat_rows = dcat.view(items * (S + I), E)
        dimf = torch.empty((items * I, E), dtype=bf16, device=dcat.device)
        ops.rows_copy(dcat_rows, S + I, S, dimf, I, 0, items, I, E)
        dtf, dimf = xit_backward(W["xit"], ctx["c_x"], dcat_rows, sink, dy_extra=dimf)
        self._bwd_calls += 1
        if self.on_trunk_grads is not None and self._bwd_calls >= self.fc1_passes:
            # every gradient except those of text_proj / img_proj is final (last backward pass of this optimizer
            # step): data parallel, their all-reduce starts here and runs under the projections' backward
            self.on_trunk_grads()
        mlp_backward(W["tp1"], W["tp2"], ctx["c_tp"], dtf, sink)
        mlp_backward(W["ip1"], W["ip2"], ctx["c_ip"], dimf, sink)
        if deferred_fc1 is not None:
            dy1p, handle = deferred_fc1
            if handle is not None:
                dy_ = handle()
                x_ = ctx["cat_all"]() if ctx.get("cat_all") is not None else self.dp_gather(ctx["cat"])
            elif self.dp_gather is not None:
                dy_, x_ = self.dp_gather(dy1p), self.dp_gather(ctx["cat"])
            else:
                dy_, x_ = dy1p, ctx["cat"]
            rows = getattr(self, "fc1_rows", None)
            if rows is not None and handle is not None:
                # row-sharded fc1 (dist.GradSync): only this rank's rows of the global-batch gradient
                r0, r1 = rows
                self._fc1_wgrad(dy_[:, r0:r1], x_, self.fc1_grad_bf16[r0:r1])
            else:
                self._fc1_wgrad(dy_, x_, self.fc1_grad_bf16)


def _to_items(src, index):
    """[bs, T_src, row] -> bf16 [bs, T_dst, row]: fp32 sources (batches as the loader yields them) go through the fused
    gather + cast kernel; bf16 sources (ppo.RolloutMemory keeps the stored rollout batches in bf16) are gathered as
    they are, or used in place when no gather is asked for."""
    if src.dtype == bf16:
        return src if index is None else ops.gather_rows(src.contiguous(), index)
    return ops.cast_gather(src, index)


def _add_bf16(a, b):
    """a + b for two bf16 [rows, D] tensors through the grouped row-copy kernel (accumulate mode)."""
    out = a.clone()
    rows, D = a.shape
    ops.rows_copy(b, rows, 0, out, rows, 0, 1, rows, D, accumulate=True)
    return out


def _small_linear(x, lin, bank):
    """n_out > 1 heads (mode 'cls': 768 -> 3): one rowdot per output."""
    outs = [ops.rowdot_fwd(x, lin.weight.detach()[j].contiguous(), lin.bias.detach()[j:j + 1].contiguous(),
                           x.shape[0]) for j in range(lin.weight.shape[0])]
    return torch.stack(outs, dim=1)


def _small_linear_bwd(x, lin, bank, dlogits, sink):
    n_out = lin.weight.shape[0]
    dx = None
    gw = torch.empty_like(lin.weight)
    gb = torch.empty_like(lin.bias)
    for j in range(n_out):
        dxj, dw, dbias = ops.rowdot_bwd(x, lin.weight.detach()[j].contiguous(), dlogits[:, j].contiguous(), x.shape[0])
        gw[j] = dw
        gb[j] = dbias[0]
        dx = dxj if dx is None else _add_bf16(dx, dxj)
    sink.put_vec(lin.weight, gw); sink.put_vec(lin.bias, gb)
    return dx
