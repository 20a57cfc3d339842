"""Data-parallel gradient synchronisation over NCCL (one process per GPU, NVLink 5 / NVSwitch).

The reference trains independent replicas in stage 2/3 (no DDP wrap: SURVEY.md §0 fact 5); the north_star adds
gradient averaging.  Design for the LR2PPO fusion model, where one matrix (out_layer.fc1, 500 M params) holds
97 % of the parameters but only sees `items` (48) activation rows per rank:

  * out_layer.fc1.weight: instead of all-reducing its 2 GB fp32 gradient, every rank ALL-GATHERS the two
    activation operands of the weight gradient (dY [items,3072] and X [items,162816], bf16: 15.9 MB per rank)
    and computes the global-batch gradient locally with a K = world*items GEMM.  NVLink traffic drops from
    2 x 2 GB to world x 16 MB per model per step.
  * everything else (19 M params): one flat fp32 bucket, SUM all-reduce.
  * the 1/world factor is folded into the fused AdamW (`grad_scale`), so no extra pass touches the gradients.
"""
import torch
import torch.distributed as dist


class Fc1Parallel:
    """out_layer.fc1 split along its INPUT dimension (K-split) over the data-parallel ranks; used by
    engine.FusionEngine when attached by GradSync.attach(..., tensor_parallel=True).

    The row-sharded optimizer of round 1 kept a complete bf16 copy of the 3072 x 162816 weight on every rank and
    re-assembled it after every step: an all-gather of 1 GB per model per step (2 x 875 MB received per rank at 8
    GPUs), the largest exposed item of the 8-GPU step.  Here rank r owns -- fp32 master, Adam moments, bf16 copy --
    the column block W[:, k0:k1) and nothing else of the weight is ever needed on it:

      forward   every rank sends column block q of its activation rows X [items, 162816] to rank q (all-to-all,
                (world-1)/world of 15.6 MB per rank); rank r multiplies everybody's block r by W[:, k0:k1)^T and the
                fp32 partial pre-activations [world*items, 3072] are summed and scattered (reduce-scatter, 4.7 MB);
                bias + GELU run on the own rows (lr2_bias_gelu_rows).
      backward  dY rows are all-gathered (0.3 MB per rank); rank r computes dX[:, k0:k1) for everybody's items --
                complete sums over the 3072 hidden units, no reduction -- and the all-to-all returns each rank its
                rows; the weight gradient of the owned block is dY_all^T X_all[:, k0:k1), local.

    Per model and step at 8 GPUs a rank exchanges ~45 MB instead of receiving ~1 GB, streams 1/8 of the weight from
    HBM in every forward / dgrad / wgrad pass and updates 1/8 of the 500 M parameters.  (A row split would need the
    whole X of every rank on every rank: an all-gather of 109 MB per forward; measured 7.83 ms/step at 8 GPUs against
    8.41 ms with the shadow all-gather, both in profiles/r02_multi_gpu.md.)  `active = False` (GradSync.replicated)
    restores replicated execution for code that the ranks do not run in lock-step (evaluation)."""

    def __init__(self, world, rank, cols, group=None):
        self.world, self.rank, self.cols, self.group, self.active = world, rank, tuple(cols), group, True

    def all_gather(self, t):
        out = torch.empty((self.world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t.contiguous(), group=self.group)
        return out

    def _all_to_all(self, t):
        out = torch.empty_like(t)
        dist.all_to_all_single(out, t, group=self.group)
        return out

    def scatter_cols(self, x):
        """x [items, world * Kb] -> [world * items, Kb]: block r of EVERY rank's rows (rank-major), on rank r."""
        items = x.shape[0]
        kb = x.shape[1] // self.world
        send = x.view(items, self.world, kb).permute(1, 0, 2).contiguous()          # [dest, items, Kb]
        return self._all_to_all(send).view(self.world * items, kb)

    def gather_cols_async(self, t):
        """Inverse of scatter_cols: t [world * items, Kb] (block `rank` of everybody's rows) -> handle; handle()
        waits and returns [items, world * Kb], this rank's rows with every column block."""
        items = t.shape[0] // self.world
        kb = t.shape[1]
        t = t.contiguous()
        out = torch.empty_like(t)
        work = dist.all_to_all_single(out, t, group=self.group, async_op=True)

        def handle():
            work.wait()
            return out.view(self.world, items, kb).permute(1, 0, 2).reshape(items, self.world * kb)
        return handle

    def reduce_scatter(self, t):
        """t [world * rows, D] -> this rank's [rows, D] block of the sum over the ranks."""
        rows = t.shape[0] // self.world
        if dist.get_backend(self.group) == "gloo":        # gloo has no reduce-scatter (CPU algebra tests): sum, slice
            full = t.contiguous().clone()
            dist.all_reduce(full, group=self.group)
            return full[self.rank * rows:(self.rank + 1) * rows].clone()
        out = torch.empty((rows,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.reduce_scatter_tensor(out, t.contiguous(), group=self.group)
        return out


class GradSync:
    def __init__(self, world, group=None):
        self.world = world
        self.group = group
        self._flat = {}
        self._skip = set()
        self._sharded = {}
        self._dirty = set()       # ids of modules whose non-owned fp32 master rows are stale on this rank
        self._dirty_opt = set()   # ids of optimizers whose non-owned Adam moment rows are stale on this rank
        self._opt_of = {}         # id(module) -> optimizer that row-shards it
        self._guarded = set()

    def broadcast_params(self, module):
        """Replicas must start identical once gradients are averaged (rank 0's initialisation wins)."""
        for p in module.parameters():
            dist.broadcast(p.data, 0, group=self.group)
        # `.data` writes do not bump version counters: mark every bf16 copy stale.  The bank object and its tensors
        # are kept (they may already be registered with an optimizer); the re-cast lands in them in place.
        for m in module.modules():
            for bank in (getattr(getattr(m, "_engine", None), "bank", None), getattr(m, "_bank", None)):
                if bank is not None:
                    bank.invalidate()

    @staticmethod
    def _engines(module):
        eng = getattr(module, "_engine", None)
        return [eng] if eng is not None else [m._engine for m in module.children() if hasattr(m, "_engine")]

    def gather_rows(self, t):
        """[rows, D] -> [world*rows, D] (all ranks' rows, rank-major)."""
        out = torch.empty((self.world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t.contiguous(), group=self.group)
        return out

    def gather_rows_async(self, t):
        """Start the all-gather of [rows, D] on NCCL's own stream and return a handle; `handle()` makes the current
        stream wait and yields the [world*rows, D] tensor.  Used to prefetch the out_layer.fc1 wgrad operand X (the
        concat buffer, 15.6 MB per rank) during the forward pass, so the exchange overlaps the rest of forward and
        backward instead of sitting in front of the wgrad GEMM."""
        out = torch.empty((self.world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        work = dist.all_gather_into_tensor(out, t.contiguous(), group=self.group, async_op=True)

        def handle():
            work.wait()
            return out
        return handle

    def attach(self, module, optimizer, shard_fc1=None, tensor_parallel=None):
        """Enable the activation-gather path for the module's out_layer.fc1 and fold 1/world into AdamW.

        shard_fc1 (default on, LR2_DP_SHARD=0 disables): ZeRO-1 style row sharding of out_layer.fc1 (97 % of the
        parameters).  Rank r computes the global-batch weight gradient only for its 1/world of the rows (K = world *
        items GEMM on a row slice), runs AdamW only on those rows (fp32 master, moments and bf16 shadow slices), and the
        updated bf16 shadow rows are all-gathered in place (`after_step`).  Every element is still updated exactly
        once per step with the same global gradient, so the arithmetic is unchanged; per rank the 14.5 GB optimizer
        pass and the redundant full-size wgrad shrink by world.  The fp32 master / moment rows of other ranks go
        stale locally: call `consolidate(module, optimizer)` before saving a checkpoint."""
        import os
        if shard_fc1 is None:
            shard_fc1 = os.environ.get("LR2_DP_SHARD", "1") == "1"
        if tensor_parallel is None:
            # tensor_parallel (default, LR2_DP_TP=0 selects round 1's row sharding + shadow all-gather): out_layer.fc1
            # split along its input dimension (Fc1Parallel above): the optimizer owns a COLUMN block per rank
            tensor_parallel = os.environ.get("LR2_DP_TP", "1") == "1"
        rank = dist.get_rank(self.group)
        early = os.environ.get("LR2_DP_EARLY_REDUCE", "1") == "1"
        for e in self._engines(module):
            e.dp_gather = self.gather_rows
            e.dp_gather_async = self.gather_rows_async
            if early and hasattr(e, "on_trunk_grads"):
                e.on_trunk_grads = (lambda m=module: self._early_reduce(m))
            w = e.m.out_layer.fc1.weight
            self._skip.add(id(w))
            e.fc1_rows = None
            rows = w.shape[0] // self.world
            kb = w.shape[1] // self.world
            can_shard = shard_fc1 and self.world > 1 and getattr(e, "fc1_grad_bf16", None) is not None
            e.tp = None
            if can_shard and tensor_parallel and w.shape[1] % self.world == 0 and kb % 128 == 0 and \
                    hasattr(optimizer, "set_col_window"):
                e.tp = Fc1Parallel(self.world, rank, (rank * kb, (rank + 1) * kb), self.group)
                optimizer.set_col_window(w, rank * kb, (rank + 1) * kb)
                self._sharded[id(module)] = (e, w)
            elif can_shard and w.shape[0] % self.world == 0 and (rows * w.shape[1]) % 4096 == 0 \
                    and hasattr(optimizer, "set_window"):
                e.fc1_rows = (rank * rows, (rank + 1) * rows)
                optimizer.set_window(w, rank, self.world)
                self._sharded[id(module)] = (e, w)
                if id(module) not in self._guarded:
                    # state_dict() of a module whose foreign rows are stale would silently save torn weights:
                    # refuse until consolidate() ran (checkpoint.save_sharded reads only the owned rows and opts out)
                    module.register_state_dict_pre_hook(self._refuse_stale_state_dict)
                    self._guarded.add(id(module))
                self._opt_of[id(module)] = optimizer
                if id(optimizer) not in self._guarded and hasattr(optimizer, "register_state_dict_pre_hook"):
                    optimizer.register_state_dict_pre_hook(self._refuse_stale_optimizer_state)
                    self._guarded.add(id(optimizer))
        optimizer.grad_scale = 1.0 / self.world
        optimizer._hyper.clear()

    def _refuse_stale_state_dict(self, module, prefix, keep_vars):
        if id(module) in self._dirty and not getattr(module, "_lr2_sharded_save", False):
            raise RuntimeError(
                "state_dict() of a module with a row-sharded out_layer.fc1: the fp32 master rows owned by other ranks "
                "are stale on this rank.  Call GradSync.consolidate(module, optimizer) first, or save with "
                "checkpoint.save_sharded(..., row_shards=GradSync.row_shards(module)).")

    def _refuse_stale_optimizer_state(self, optimizer):
        if id(optimizer) in self._dirty_opt and not getattr(optimizer, "_lr2_sharded_save", False):
            raise RuntimeError(
                "state_dict() of an optimizer that row-shards out_layer.fc1: the Adam moments of rows owned by other "
                "ranks are stale on this rank.  Call GradSync.consolidate(module, optimizer) first, or save with "
                "checkpoint.save_sharded.")

    def after_step(self, module):
        """After optimizer.step(): start the in-place all-gather of the updated bf16 shadow rows of a row-sharded
        out_layer.fc1 on NCCL's stream.  Returns wait(); it must be called before the module's next forward."""
        ent = self._sharded.get(id(module))
        if ent is None:
            return lambda: None
        e, w = ent
        self._dirty.add(id(module))
        if id(module) in self._opt_of:
            self._dirty_opt.add(id(self._opt_of[id(module)]))
        if e.tp is not None:
            return lambda: None              # K-split execution never reads foreign column blocks: nothing to gather
        shadow = e.bank.get(w)
        r0, r1 = e.fc1_rows
        work = dist.all_gather_into_tensor(shadow, shadow[r0:r1], group=self.group, async_op=True)
        return work.wait

    def gather_shadow(self, module):
        """All-gather the bf16 rows of a row-parallel out_layer.fc1 so that every rank holds the complete weight copy
        again (blocking).  Needed before code that does not run in lock-step on all ranks."""
        ent = self._sharded.get(id(module))
        if ent is None:
            return
        e, w = ent
        shadow = e.bank.get(w)
        if e.tp is not None:
            self._gather_cols(shadow, e.tp.cols)
            return
        r0, r1 = e.fc1_rows
        dist.all_gather_into_tensor(shadow, shadow[r0:r1].clone(), group=self.group)

    def _gather_cols(self, t, cols):
        """Complete the 2-D tensor t (every rank holds valid data in its own column block) on all ranks, in place."""
        k0, k1 = cols
        pieces = torch.empty((self.world, t.shape[0], k1 - k0), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(pieces, t[:, k0:k1].contiguous(), group=self.group)
        for q in range(self.world):
            t[:, q * (k1 - k0):(q + 1) * (k1 - k0)].copy_(pieces[q])

    def replicated(self, *modules):
        """Context manager: inside it the given modules run replicated (complete bf16 weights, no collectives in
        forward) -- e.g. evaluation, where ranks score different numbers of clips per forward."""
        import contextlib

        @contextlib.contextmanager
        def scope():
            tps = []
            for m in modules:
                ent = self._sharded.get(id(m))
                if ent is not None and ent[0].tp is not None:
                    self.gather_shadow(m)
                    ent[0].tp.active = False
                    tps.append(ent[0].tp)
            try:
                yield
            finally:
                for tp in tps:
                    tp.active = True
        return scope()

    def row_shards(self, module):
        """{parameter name: (r0, r1)} of the rows of `module`'s row-sharded parameters that this rank owns
        ({} when nothing is sharded): the `row_shards` argument of checkpoint.save_sharded, which then needs no
        consolidate()."""
        ent = self._sharded.get(id(module))
        if ent is None:
            return {}
        e, w = ent
        for name, p in module.named_parameters():
            if p is w:
                # (r0, r1) = a row block; (dim, lo, hi) = a block along `dim` (K-split: columns)
                return {name: (1,) + tuple(e.tp.cols) if e.tp is not None else tuple(e.fc1_rows)}
        return {}

    def consolidate(self, module, optimizer=None):
        """Make the fp32 master weight (and, with `optimizer`, Adam's moments) of a row-sharded out_layer.fc1 complete
        on every rank again (in-place all-gathers); call before state_dict() / checkpoint.save_*."""
        ent = self._sharded.get(id(module))
        if ent is None:
            return
        e, w = ent
        tensors = [w.data]
        if optimizer is not None:
            st = optimizer.state_for(w)
            tensors += [st["exp_avg"], st["exp_avg_sq"]]
        for t in tensors:
            if e.tp is not None:
                self._gather_cols(t, e.tp.cols)
            else:
                r0, r1 = e.fc1_rows
                dist.all_gather_into_tensor(t, t[r0:r1], group=self.group)
        self._dirty.discard(id(module))          # master weight complete again: module.state_dict() is safe
        if optimizer is not None:
            self._dirty_opt.discard(id(optimizer))

    def start(self, module):
        """Begin the SUM all-reduce of every gradient except out_layer.fc1 (whose gradient is already global) on
        NCCL's own stream; returns `finish()`, which waits for it.  FusedAdamW.step(first={fc1}, between=finish)
        updates the 500 M fc1 parameters meanwhile.

        The gradients live in ONE flat fp32 bucket (16-byte aligned slots).  With persistent gradient buffers
        (engine.persistent_grads, used by the CUDA-graph step) the `.grad` tensors are re-pointed to views of the
        bucket on first use, so later steps all-reduce in place with no copy in or out; otherwise the values are
        copied in before and back after the collective."""
        params = [p for p in module.parameters() if p.grad is not None and id(p) not in self._skip]
        if not params:
            return lambda: None
        # bucket order: first everything whose gradient is final before the backward reaches the input projections
        # (head, xitt, out_layer, xit -- the "trunk"), then text_proj / img_proj; see _early_reduce
        late = {id(p) for name in ("text_proj", "img_proj") if hasattr(module, name)
                for p in getattr(module, name).parameters()}
        params = [p for p in params if id(p) not in late] + [p for p in params if id(p) in late]
        grads = [p.grad for p in params]
        ent = self._flat.get(id(module))
        if ent is None or ent["shapes"] != [tuple(g.shape) for g in grads]:
            offs, total, split = [], 0, None
            for p, g in zip(params, grads):
                if split is None and id(p) in late:
                    split = total
                offs.append(total)
                total += (g.numel() + 3) // 4 * 4
            flat = torch.zeros(total, dtype=torch.float32, device=grads[0].device)
            ent = {"flat": flat, "shapes": [tuple(g.shape) for g in grads], "split": split or 0, "inplace": False,
                   "early": None, "views": [flat[o:o + g.numel()] for o, g in zip(offs, grads)]}
            self._flat[id(module)] = ent
        flat, views = ent["flat"], ent["views"]
        inplace = all(g.data_ptr() == v.data_ptr() for g, v in zip(grads, views))
        copy_back = False
        if not inplace:
            if ent["early"] is not None:                     # cannot happen: _early_reduce needs the in-place state
                ent["early"].wait(); ent["early"] = None
            torch._foreach_copy_(views, [g.reshape(-1) for g in grads])
            if all(getattr(e, "persistent_grads", False) for e in self._engines(module)):
                for p, v in zip(params, views):
                    p.grad = v.view(p.shape)             # from now on the backward writes straight into the bucket
                inplace = True
            else:
                copy_back = True
        ent["inplace"] = inplace
        early, ent["early"] = ent["early"], None
        # the trunk part may already be in flight (started inside backward); then only the projections are left
        work = dist.all_reduce(flat[ent["split"]:] if early is not None else flat, group=self.group, async_op=True)

        def finish():
            if early is not None:
                early.wait()
            work.wait()
            if copy_back:
                torch._foreach_copy_([g.view(-1) for g in grads], views)
        return finish

    def _early_reduce(self, module):
        """engine.on_trunk_grads: called inside backward when every gradient of the bucket's first part is final.
        Only in the zero-copy state (persistent gradient buffers that ARE views of the bucket): the all-reduce of that
        part (about half of the 19 M small parameters) then runs on NCCL's stream under the backward of the input
        projections instead of after the whole backward.  LR2_DP_EARLY_REDUCE=0 disables it."""
        ent = self._flat.get(id(module))
        if ent is None or not ent["inplace"] or not ent["split"] or ent["early"] is not None:
            return
        ent["early"] = dist.all_reduce(ent["flat"][:ent["split"]], group=self.group, async_op=True)

    def early_params(self, module):
        """ids of the parameters whose gradients need no all-reduce (out_layer.fc1 of each engine)."""
        return {i for i in self._skip if any(id(p) == i for p in module.parameters())}

    def __call__(self, module):
        self.start(module)()
