"""AdamW + LR schedules with the reference's API (tencentpretrain/utils/optimizers.py:62-86, 305-402),
running as ONE multi-tensor CUDA launch per parameter group (lr2_adamw_multi).

`FusedAdamW(params, lr, betas, eps, weight_decay, correct_bias)` takes the same arguments and
param-group dicts as the reference AdamW, is a torch.optim.Optimizer (so LambdaLR schedulers work
unchanged), keeps `exp_avg` / `exp_avg_sq` per parameter, and can refresh bf16 shadow copies of the
weights (read by the GEMMs) in the same pass.
"""
import math
import struct

import torch
from torch.optim import Optimizer
from torch.optim.lr_scheduler import LambdaLR

from . import _lib
from ._lib import check


def _f2i(x):
    return struct.unpack("<i", struct.pack("<f", float(x)))[0]


def subtract_span(spans, cut):
    """Chunk-index bookkeeping of the two-phase / row-sharded step: `spans` (list of (start, count)) minus the
    interval cut = (start, count)."""
    ca, cb = cut[0], cut[0] + cut[1]
    out = []
    for a, n in spans:
        b = a + n
        if cb <= a or ca >= b:
            out.append((a, n))
            continue
        if a < ca:
            out.append((a, ca - a))
        if cb < b:
            out.append((cb, b - cb))
    return out


class FusedAdamW(Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0, correct_bias=True,
                 shadow_bf16=False, grad_scale=1.0):
        if lr < 0.0:
            raise ValueError("Invalid learning rate: {} - should be >= 0.0".format(lr))
        if not 0.0 <= betas[0] < 1.0:
            raise ValueError("Invalid beta parameter: {} - should be in [0.0, 1.0[".format(betas[0]))
        if not 0.0 <= betas[1] < 1.0:
            raise ValueError("Invalid beta parameter: {} - should be in [0.0, 1.0[".format(betas[1]))
        if not 0.0 <= eps:
            raise ValueError("Invalid epsilon value: {} - should be >= 0.0".format(eps))
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, correct_bias=correct_bias)
        super().__init__(params, defaults)
        self._shadow = shadow_bf16
        self.grad_scale = grad_scale
        self._tables = {}      # group index -> dict
        self._shadows = {}     # id(p) -> bf16 tensor (may be registered externally)
        self._grad_src = {}    # id(p) -> tensor used as gradient instead of p.grad (e.g. bf16 wgrad buffers)
        self._fused = {}       # id(p) -> provider() of [(dY, X)] for the fused wgrad+AdamW kernel
        self._hyper = {}       # group index -> device hyper-parameter buffer
        self.frozen_hyper = False  # True while the step is replayed from a CUDA graph
        self.launch_bytes = []     # algorithmic HBM bytes of every lr2_adamw_multi launch made while _lib.PROFILE is on

    # -- extras -----------------------------------------------------------
    def register_shadow(self, p, shadow):
        """Use `shadow` (bf16, same shape) as the low-precision copy refreshed at every step."""
        self._shadows[id(p)] = shadow
        self._tables.clear()

    def register_grad(self, p, grad):
        """Read the gradient of `p` from `grad` (fp32 or bf16 buffer owned by the caller)."""
        self._grad_src[id(p)] = grad
        self._tables.clear()

    def register_fused_wgrad(self, p, provider):
        """Update the 2-D weight `p` with lr2_gemm_wgrad_adamw: `provider()` returns the list of
        (dY [rows, out] bf16, X [rows, in] bf16) activations stashed by backward since the last step; the
        gradient dY^T X is consumed inside the kernel and never materialised (p.grad stays None)."""
        self._fused[id(p)] = provider
        self._tables.clear()

    def set_window(self, p, k, n):
        """Data-parallel row sharding (dist.GradSync): this rank updates only the k-th of n equal chunk ranges of
        parameter p (out_layer.fc1: 1/n of its rows); the other chunks are never launched here."""
        if not hasattr(self, "_windows"):
            self._windows = {}
        self._windows[id(p)] = (int(k), int(n))
        self._tables.clear()

    def set_col_window(self, p, c0, c1):
        """Data-parallel column sharding (dist.GradSync, K-split out_layer.fc1): this rank updates only columns
        [c0, c1) of the 2-D parameter p; nothing else of p is ever launched here."""
        if not hasattr(self, "_col_windows"):
            self._col_windows = {}
        self._col_windows[id(p)] = (int(c0), int(c1))
        self._tables.clear()

    def _owned(self, tab, i):
        """Chunk range (start, count) of tensor id i that this rank updates."""
        a, n = tab["ranges"][i]
        w = getattr(self, "_windows", {}).get(i)
        if w is None:
            return a, n
        k, parts = w
        if n % parts:
            raise _lib.Lr2Error("sharded parameter: chunk count is not divisible by the number of ranks")
        return a + k * (n // parts), n // parts

    def shadow_of(self, p):
        return self._shadows.get(id(p))

    def state_for(self, p):
        return self.state[p]

    # -- internals --------------------------------------------------------
    def _grad_of(self, p):
        g = self._grad_src.get(id(p))
        return g if g is not None else p.grad

    def _build(self, gi, group, ps):
        L = _lib.load()
        chunk = L.lr2_adamw_chunk_elems()
        dev = ps[0].device
        ptrs, meta, chunks, gptrs, lens = [], [], [], [], []
        for t, p in enumerate(ps):
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                raise _lib.Lr2Error("FusedAdamW needs contiguous fp32 CUDA parameters (no CPU fallback)")
            st = self.state[p]
            if len(st) == 0:
                st["step"] = 0
                st["exp_avg"] = torch.zeros_like(p.data)
                st["exp_avg_sq"] = torch.zeros_like(p.data)
            sh = self._shadows.get(id(p))
            if sh is None and self._shadow:
                sh = p.data.to(torch.bfloat16)
                self._shadows[id(p)] = sh
            g = self._grad_of(p)
            if g.dtype not in (torch.float32, torch.bfloat16) or not g.is_contiguous() or g.numel() != p.numel():
                raise _lib.Lr2Error("gradient must be contiguous fp32/bf16 of the parameter's size")
            ptrs += [p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                     sh.data_ptr() if sh is not None else 0, 0]
            meta += [p.numel(), _f2i(group["weight_decay"]), int(g.dtype == torch.bfloat16), 0]
            gptrs.append(g.data_ptr())
            cw = getattr(self, "_col_windows", {}).get(id(p))
            if cw is None:
                offs = torch.arange(0, p.numel(), chunk, dtype=torch.int64)
                lens.append(torch.clamp(p.numel() - offs, max=chunk))
                chunks.append(torch.stack([torch.full_like(offs, t), offs], dim=1))
            else:
                # column-sharded 2-D parameter: this rank owns columns [c0, c1) of every row; each row segment is
                # cut into pieces of at most `chunk` elements with explicit lengths (multiples of 4)
                c0, c1 = cw
                rows, cols = p.shape
                if (c1 - c0) % 4 or c0 % 4 or cols % 4:
                    raise _lib.Lr2Error("column window must be aligned to 4 elements")
                seg = torch.arange(0, c1 - c0, chunk, dtype=torch.int64)
                seg_len = torch.clamp((c1 - c0) - seg, max=chunk)
                offs = (torch.arange(rows, dtype=torch.int64)[:, None] * cols + c0 + seg[None, :]).reshape(-1)
                ln = seg_len[None, :].expand(rows, -1).reshape(-1)
                lens.append(ln)
                chunks.append(torch.stack([torch.full_like(offs, t) | (ln << 32), offs], dim=1))
        chunk_t = torch.cat(chunks, dim=0).contiguous()
        starts, acc = {}, 0
        for t, p in enumerate(ps):          # chunk range of every tensor (two-phase step: see step(first=...))
            starts[id(p)] = (acc, chunks[t].shape[0])
            acc += chunks[t].shape[0]
        # bytes moved by chunks [0, i): lets a profiling pass attribute algorithmic bytes to each launch (chunk span)
        per_chunk = torch.cat([lens[t] * self._bytes_per_element(p, self._grad_of(p)) for t, p in enumerate(ps)])
        prefix = torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(per_chunk, 0)])
        tab = dict(params=ps, ids=[id(p) for p in ps], gptrs=gptrs, n_chunks=chunk_t.shape[0], ranges=starts,
                   byte_prefix=prefix,
                   ptrs=torch.tensor(ptrs, dtype=torch.int64, device=dev),
                   meta=torch.tensor(meta, dtype=torch.int64, device=dev),
                   chunks=chunk_t.to(dev))
        return tab

    def _refresh_ptrs(self, tab):
        """Only the gradient addresses changed (autograd / zero_grad re-allocated them): patch the pointer table."""
        ptrs = tab["ptrs"].cpu()
        gptrs = []
        for t, p in enumerate(tab["params"]):
            g = self._grad_of(p)
            ptrs[6 * t + 1] = g.data_ptr()
            gptrs.append(g.data_ptr())
        tab["ptrs"].copy_(ptrs)
        tab["gptrs"] = gptrs

    def _group_t(self, group):
        """Bias-correction step count of the NEXT update of this group = state[p]['step'] + 1 of its first parameter
        that has state (the reference keeps it per parameter, tencentpretrain/utils/optimizers.py:379; all
        parameters of a group step together here).  It lives in the optimizer state, so state_dict() saves it and a
        resume / GradSync.attach does not restart the correction."""
        for p in group["params"]:
            st = self.state.get(p)
            if st and "step" in st:
                return int(st["step"]) + 1
        return 1

    def _hyper_values(self, group, t):
        beta1, beta2 = group["betas"]
        lr = group["lr"]
        step_size = lr
        if group["correct_bias"]:
            step_size = lr * math.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t)
        return (step_size, beta1, beta2, group["eps"], 1.0 - beta1, 1.0 - beta2, self.grad_scale, lr)

    def _hyper_buf(self, gi, group, dev):
        """Device-side hyper-parameters of group gi: {step_size, b1, b2, eps, 1-b1, 1-b2, grad_scale, lr}."""
        ent = self._hyper.get(gi)
        if ent is None:
            ent = {"dev": torch.zeros(8, dtype=torch.float32, device=dev), "host": None}
            self._hyper[gi] = ent
        if self.frozen_hyper and ent["host"] is not None:
            return ent["dev"]          # inside a captured CUDA graph: the caller refreshes it with update_hyper()
        hv = self._hyper_values(group, self._group_t(group))
        if ent["host"] != hv:
            ent["dev"].copy_(torch.tensor(hv, dtype=torch.float32))
            ent["host"] = hv
        return ent["dev"]

    def algorithmic_bytes(self, gi):
        """HBM bytes one lr2_adamw_multi launch of group gi must move: read p, g, m, v + write p, m, v (+ bf16 shadow)."""
        total = 0
        for p in self.param_groups[gi]["params"]:
            g = self._grad_of(p)
            if g is None or id(p) in self._fused:
                continue
            parts = getattr(self, "_windows", {}).get(id(p), (0, 1))[1]      # row-sharded: this rank moves 1/parts
            n = p.numel() // parts
            cw = getattr(self, "_col_windows", {}).get(id(p))
            if cw is not None:
                n = p.shape[0] * (cw[1] - cw[0])                             # column-sharded
            total += n * self._bytes_per_element(p, g)
        return total

    def _bytes_per_element(self, p, g):
        return 4 + g.element_size() + 4 + 4 + 4 + 4 + 4 + (2 if id(p) in self._shadows else 0)

    _PIN_SLOTS = 4

    def update_hyper(self):
        """Recompute lr / bias-correction on the host and upload the 8 floats of every group asynchronously.  Used with
        CUDA graphs, where `step()` itself must not copy from the host (and is not re-run per replay: this call also
        advances the per-parameter step counts the replay stands for).  The pinned staging is a small ring, each slot
        guarded by an event recorded after its copy, so a host that runs ahead of the replays never overwrites a
        buffer whose H2D copy has not executed yet."""
        for gi, group in enumerate(self.param_groups):
            ent = self._hyper.get(gi)
            if ent is None:
                continue
            hv = self._hyper_values(group, self._group_t(group))
            for p in group["params"]:
                st = self.state.get(p)
                if st and "step" in st:
                    st["step"] += 1
            if ent["host"] != hv:
                ring = ent.setdefault("ring", [])
                k = ent.get("ring_i", 0)
                if len(ring) <= k:
                    ring.append((torch.empty(8, dtype=torch.float32).pin_memory(), torch.cuda.Event()))
                else:
                    ring[k][1].synchronize()          # the copy that last used this slot has executed
                pin, ev = ring[k]
                pin.copy_(torch.tensor(hv, dtype=torch.float32))
                ent["dev"].copy_(pin, non_blocking=True)
                ev.record(torch.cuda.current_stream(ent["dev"].device))
                ent["ring_i"] = (k + 1) % self._PIN_SLOTS
                ent["host"] = hv

    @torch.no_grad()
    def step(self, closure=None, first=None, between=None):
        """first / between (data-parallel overlap): the chunks of the parameters whose ids are in `first` (the
        out_layer.fc1 weight, whose gradient is already global) are launched before `between()` is called (it waits
        for the asynchronous all-reduce of the remaining gradients), everything else after it.  Per-element
        arithmetic and results are identical to a single launch."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        L = _lib.load()
        prepared = []
        for gi, group in enumerate(self.param_groups):
            live = [p for p in group["params"] if id(p) not in self._fused and self._grad_of(p) is not None]
            fused = [p for p in group["params"] if id(p) in self._fused]
            if not live and not fused:
                continue
            hyper = self._hyper_buf(gi, group, (live or fused)[0].device)
            tab = self._tables.get(gi)
            if tab is not None and [id(p) for p in live] != tab["ids"]:
                tab = None
            if tab is not None and [self._grad_of(p).data_ptr() for p in live] != tab["gptrs"]:
                self._refresh_ptrs(tab)
            if tab is None and live:
                tab = self._build(gi, group, live)
                self._tables[gi] = tab
            spans = []
            if live:
                if not (live[0].is_cuda and torch.cuda.is_current_stream_capturing()):
                    for p in live:                       # a capture pass executes nothing: update_hyper() counts the
                        self.state[p]["step"] += 1       # steps its replays stand for
                spans = [(0, tab["n_chunks"])]
            prepared.append((group, live, fused, hyper, tab, spans))

        def launch(tab, hyper, a, n):
            if _lib.PROFILE is not None:         # bench.py's roofline pass: algorithmic bytes of this launch
                self.launch_bytes.append(int(tab["byte_prefix"][a + n] - tab["byte_prefix"][a]))
            _lib.run(L.lr2_adamw_multi, tab["ptrs"].data_ptr(), tab["meta"].data_ptr(),
                     tab["chunks"].data_ptr() + 16 * a, n, hyper.data_ptr(), _lib.stream())

        subtract = subtract_span

        windows = getattr(self, "_windows", {})
        for k, (group, live, fused, hyper, tab, spans) in enumerate(prepared):
            if not live:
                continue
            # chunks of row-sharded tensors that belong to other ranks are never launched
            for i in windows:
                if i in tab["ranges"]:
                    full = tab["ranges"][i]
                    own = self._owned(tab, i)
                    spans = subtract(spans, full) + [own]
            early = sorted(self._owned(tab, i) for i in (first or ()) if i in tab["ranges"])
            if early and between is not None:
                # phase 1: the early tensors of every group; phase 2 (after between()): all remaining chunk spans
                for a, n in early:
                    launch(tab, hyper, a, n)
                    spans = subtract(spans, (a, n))
            prepared[k] = (group, live, fused, hyper, tab, sorted(spans))
        if between is not None:
            between()
        stale = []
        for group, live, fused, hyper, tab, spans in prepared:
            for a, n in spans:
                launch(tab, hyper, a, n)
            # the kernel wrote these masters through raw pointers: bump their version counters so that any bf16 copy
            # NOT refreshed by this pass (engine.ShadowBank entries that were never registered here) is re-cast at
            # its next use instead of silently freezing the forward weights.  Registered shadows are already fresh.
            stale += [p for p in live if id(p) not in self._shadows]
            for p in fused:
                pairs = self._fused[id(p)]()
                if not pairs:
                    continue
                dy = pairs[0][0] if len(pairs) == 1 else torch.cat([a for a, _ in pairs], dim=0)
                x = pairs[0][1] if len(pairs) == 1 else torch.cat([b for _, b in pairs], dim=0)
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p.data)
                    st["exp_avg_sq"] = torch.zeros_like(p.data)
                st["step"] += 1
                sh = self._shadows.get(id(p))
                out_f, in_f = p.shape
                _lib.run(L.lr2_gemm_wgrad_adamw, dy.data_ptr(), dy.stride(0), x.data_ptr(), x.stride(0), dy.shape[0],
                         out_f, in_f, p.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                         sh.data_ptr() if sh is not None else None, hyper.data_ptr(), float(group["weight_decay"]),
                         _lib.stream())
                if sh is None:
                    stale.append(p)
        if stale:
            torch._C._increment_version(stale)
        return loss


AdamW = FusedAdamW


def get_linear_schedule_with_warmup(optimizer, num_warmup_steps, num_training_steps, last_epoch=-1):
    """ref: tencentpretrain/utils/optimizers.py:62-86 (LambdaLR; lambda(0) is applied at construction)."""
    def lr_lambda(current_step):
        if current_step < num_warmup_steps:
            return float(current_step) / float(max(1, num_warmup_steps))
        return max(0.0, float(num_training_steps - current_step) /
                   float(max(1, num_training_steps - num_warmup_steps)))
    return LambdaLR(optimizer, lr_lambda, last_epoch)


def get_constant_schedule(optimizer, last_epoch=-1):
    return LambdaLR(optimizer, lambda _: 1, last_epoch=last_epoch)


def get_constant_schedule_with_warmup(optimizer, num_warmup_steps, last_epoch=-1):
    def lr_lambda(current_step):
        if current_step < num_warmup_steps:
            return float(current_step) / float(max(1.0, num_warmup_steps))
        return 1.0
    return LambdaLR(optimizer, lr_lambda, last_epoch=last_epoch)


def decay_groups(named_parameters):
    """The two parameter groups every stage script builds (finetune/ppo.py:381-393, finetune/pointwise.py:275-282):
    weight decay 0.01 except for names containing bias / gamma / beta."""
    named = list(named_parameters)
    no_decay = ["bias", "gamma", "beta"]
    return [{"params": [p for n, p in named if not any(nd in n for nd in no_decay)], "weight_decay": 0.01},
            {"params": [p for n, p in named if any(nd in n for nd in no_decay)], "weight_decay": 0.0}]


def make_scheduler(args, optimizer):
    """Scheduler selection of the stage scripts (finetune/pointwise.py:289-296, finetune/ppo.py:405-418)."""
    sched = getattr(args, "scheduler", "linear")
    if sched not in str2scheduler:
        raise ValueError(f"scheduler {sched!r} is outside the LR2PPO hot path (linear / constant / "
                         "constant_with_warmup are what the scripts can select)")
    if sched == "constant":
        return str2scheduler[sched](optimizer)
    if sched == "constant_with_warmup":
        return str2scheduler[sched](optimizer, args.train_steps * args.warmup)
    return str2scheduler[sched](optimizer, args.train_steps * args.warmup, args.train_steps)


def attach_shadows(engine, optimizer):
    """Let the optimizer refresh the engine's bf16 weight copies in its own pass (no separate cast kernels)."""
    for group in optimizer.param_groups:
        for p in group["params"]:
            if p.dim() >= 2 and p.is_cuda:
                optimizer.register_shadow(p, engine.bank.get(p))


str2optimizer = {"adamw": FusedAdamW}
str2scheduler = {"linear": get_linear_schedule_with_warmup, "constant": get_constant_schedule,
                 "constant_with_warmup": get_constant_schedule_with_warmup}
