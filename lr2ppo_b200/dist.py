"""Data-parallel gradient synchronisation over NCCL (one process per GPU, NVLink 5 / NVSwitch).

The reference trains independent replicas in stage 2/3 (no DDP wrap: SURVEY.md §0 fact 5); the
north_star adds a gradient all-reduce.  Gradients are SUM-reduced and the 1/world factor is folded
into the fused AdamW (`grad_scale`), so no extra pass touches the 2 GB out_layer.fc1 gradient.
Small gradients travel in one flat bucket; tensors >= 64 MB are reduced in place.
"""
import torch
import torch.distributed as dist

BIG = 1 << 24  # elements


class GradSync:
    def __init__(self, world, group=None):
        self.world = world
        self.group = group
        self._flat = {}

    def broadcast_params(self, module):
        """Replicas must start identical once gradients are averaged (rank 0's initialisation wins)."""
        for p in module.parameters():
            dist.broadcast(p.data, 0, group=self.group)
        eng = getattr(module, "_engine", None)
        engines = [eng] if eng is not None else [m._engine for m in module.children() if hasattr(m, "_engine")]
        for e in engines:
            e.bank = type(e.bank)()      # bf16 shadows are re-cast from the broadcast weights

    def __call__(self, module):
        small = [p.grad for p in module.parameters() if p.grad is not None and p.grad.numel() < BIG]
        big = [p.grad for p in module.parameters() if p.grad is not None and p.grad.numel() >= BIG]
        works = [dist.all_reduce(g, group=self.group, async_op=True) for g in big]
        if small:
            n = sum(g.numel() for g in small)
            flat = self._flat.get(id(module))
            if flat is None or flat.numel() != n:
                flat = torch.empty(n, dtype=torch.float32, device=small[0].device)
                self._flat[id(module)] = flat
            views = list(flat.split([g.numel() for g in small]))
            torch._foreach_copy_(views, [g.view(-1) for g in small])
            dist.all_reduce(flat, group=self.group)
            torch._foreach_copy_([g.view(-1) for g in small], views)
        for w in works:
            w.wait()
