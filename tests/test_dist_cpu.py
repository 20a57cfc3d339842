"""N>1 host logic on CPU (gloo, world_size 2): the data-parallel gradient synchronisation of lr2ppo_b200.dist.
The CUDA kernels are not involved; this checks the collective plumbing and the algebra that lets out_layer.fc1
all-gather its wgrad operands instead of all-reducing the 2 GB gradient."""
import os

import pytest
import tempfile

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class _Eng:
    def __init__(self, m):
        self.m, self.dp_gather, self.bank = m, None, type("B", (), {"invalidate": lambda self: None})()
        self.on_trunk_grads = None            # set by GradSync.attach (engine.FusionEngine.on_trunk_grads)


class _Net(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.out_layer = torch.nn.Module()
        self.out_layer.fc1 = torch.nn.Linear(24, 8)
        self.other = torch.nn.Linear(8, 4)
        self.text_proj = torch.nn.Linear(5, 3)            # an input projection: last gradients of a backward pass
        self._engine = _Eng(self)


def _worker(rank, world, path):
    from lr2ppo_b200.dist import GradSync
    dist.init_process_group("gloo", init_method=f"file://{path}", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)
    net = _Net()
    sync = GradSync(world)
    sync.broadcast_params(net)                                    # rank 0's weights everywhere
    w0 = [p.detach().clone() for p in net.parameters()]
    gathered = [torch.empty_like(w0[0]) for _ in range(world)]
    dist.all_gather(gathered, w0[0])
    assert all(torch.equal(g, gathered[0]) for g in gathered)

    class _Opt:
        grad_scale = 1.0
        _hyper = {}
    opt = _Opt()
    sync.attach(net, opt)
    assert opt.grad_scale == 1.0 / world and net._engine.dp_gather is not None
    # per-rank activations of the big layer and gradients of the small ones
    dy = torch.randn(6, 8); x = torch.randn(6, 24)
    net.other.weight.grad = torch.full_like(net.other.weight, float(rank + 1))
    net.other.bias.grad = torch.full_like(net.other.bias, float(10 * (rank + 1)))
    net.out_layer.fc1.bias.grad = dy.sum(0)
    # (1) gather path: global-batch wgrad computed locally == sum over ranks of the local wgrads
    g_dy, g_x = net._engine.dp_gather(dy), net._engine.dp_gather(x)
    assert g_dy.shape == (6 * world, 8) and torch.equal(g_dy[6 * rank:6 * rank + 6], dy)     # rank-major rows
    local = dy.t() @ x
    summed = local.clone(); dist.all_reduce(summed)
    assert torch.allclose(g_dy.t() @ g_x, summed, atol=1e-5)
    net.out_layer.fc1.weight.grad = g_dy.t() @ g_x
    before = net.out_layer.fc1.weight.grad.clone()
    # (2) bucket path: everything except fc1.weight is SUM-reduced; fc1.weight is left alone
    sync(net)
    assert torch.equal(net.out_layer.fc1.weight.grad, before)
    assert torch.allclose(net.other.weight.grad, torch.full_like(net.other.weight, sum(range(1, world + 1))))
    assert torch.allclose(net.other.bias.grad, torch.full_like(net.other.bias, 10.0 * sum(range(1, world + 1))))
    bsum = dy.sum(0).clone(); dist.all_reduce(bsum)
    assert torch.allclose(net.out_layer.fc1.bias.grad, bsum, atol=1e-5)
    # (3) persistent-gradient mode: the .grad tensors become views of the flat bucket (16-byte aligned slots) on the
    # first call; afterwards start()/finish() all-reduce in place with no copies, and fc1.weight stays untouched
    net._engine.persistent_grads = True
    small = [p for n, p in net.named_parameters() if n != "out_layer.fc1.weight"]
    for p in small:
        p.grad = torch.full_like(p, float(rank + 1))
    finish = sync.start(net)
    finish()
    ptrs = [p.grad.data_ptr() for p in small]
    base = min(ptrs)
    assert all((q - base) % 16 == 0 for q in ptrs)                          # aligned slots of one buffer
    for p in small:
        assert torch.allclose(p.grad, torch.full_like(p, float(sum(range(1, world + 1)))))
    for p in small:                                                          # "next backward" writes in place
        p.grad.fill_(float(2 * (rank + 1)))
    sync.start(net)()
    assert [p.grad.data_ptr() for p in small] == ptrs                        # still the same views: zero-copy
    for p in small:
        assert torch.allclose(p.grad, torch.full_like(p, float(2 * sum(range(1, world + 1)))))
    assert torch.equal(net.out_layer.fc1.weight.grad, before)
    assert sync.early_params(net) == {id(net.out_layer.fc1.weight)}
    # (4) early reduce: the bucket holds the trunk first, text_proj / img_proj last; the engine's hook starts the
    # all-reduce of the trunk part inside backward, start() then only reduces the projections' part
    proj = list(net.text_proj.parameters())
    assert max(p.grad.data_ptr() for p in small if all(p is not q for q in proj)) < min(q.grad.data_ptr() for q in proj)
    assert net._engine.on_trunk_grads is not None
    for p in small:
        p.grad.fill_(float(3 * (rank + 1)))
    for q in proj:
        q.grad.fill_(-1.0)                                                   # "not written yet" at hook time
    net._engine.on_trunk_grads()                                             # trunk part in flight ...
    for q in proj:
        q.grad.fill_(float(5 * (rank + 1)))                                  # ... while the projections' backward runs
    sync.start(net)()
    tot = float(sum(range(1, world + 1)))
    for p in small:
        want = 5 * tot if any(p is q for q in proj) else 3 * tot             # every slot reduced exactly once
        assert torch.allclose(p.grad, torch.full_like(p, want)), (want, p.grad.flatten()[:3])
    net._engine.on_trunk_grads(); sync.start(net)()                         # and again (state resets every step)
    for p in small:
        want = 5 * tot * world if any(p is q for q in proj) else 3 * tot * world    # equal values on all ranks now
        assert torch.allclose(p.grad, torch.full_like(p, want))
    dist.barrier()
    dist.destroy_process_group()


def test_gradsync_world2_gloo():
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(2, os.path.join(d, "rdzv")), nprocs=2, join=True)


# ---- row-sharded out_layer.fc1 optimizer: attach -> step on own rows -> shadow all-gather -> sharded checkpoint ----
class _Bank:
    def __init__(self):
        self.d = {}

    def invalidate(self):       # host stand-in of engine.ShadowBank.invalidate: re-cast in place at the next get()
        self.stale = True

    def get(self, w):
        if id(w) not in self.d:
            self.d[id(w)] = w.detach().to(torch.bfloat16)
        return self.d[id(w)]


class _Net2(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.out_layer = torch.nn.Module()
        self.out_layer.fc1 = torch.nn.Linear(1024, 8)          # rows/world * cols is a multiple of the 4096-element chunk
        self.other = torch.nn.Linear(8, 4)
        self._engine = _Eng(self)
        self._engine.bank = _Bank()
        self._engine.fc1_grad_bf16 = torch.zeros(8, 1024, dtype=torch.bfloat16)
        self._engine.dp_gather_async = None


class _RowAdam(torch.optim.Optimizer):
    """Host stand-in for FusedAdamW with the attributes GradSync uses (set_window / state_for / grad_scale)."""

    def __init__(self, params):
        super().__init__(params, dict(lr=0.1))
        self.grad_scale, self._hyper, self.windows = 1.0, {}, {}
        for p in self.param_groups[0]["params"]:
            self.state[p] = {"step": 0, "exp_avg": torch.zeros_like(p), "exp_avg_sq": torch.zeros_like(p)}

    def set_window(self, p, k, n):
        self.windows[id(p)] = (k, n)

    def state_for(self, p):
        return self.state[p]

    @torch.no_grad()
    def step(self, grads, shadows):
        for p in self.param_groups[0]["params"]:
            g = grads[id(p)] * self.grad_scale
            rows = slice(None)
            if id(p) in self.windows:
                k, n = self.windows[id(p)]
                per = p.shape[0] // n
                rows = slice(k * per, (k + 1) * per)
            st = self.state[p]
            st["step"] += 1
            st["exp_avg"][rows] = 0.9 * st["exp_avg"][rows] + 0.1 * g[rows]
            st["exp_avg_sq"][rows] = 0.999 * st["exp_avg_sq"][rows] + 0.001 * g[rows] ** 2
            p[rows] -= 0.1 * g[rows]
            if id(p) in shadows:
                shadows[id(p)][rows] = p[rows].to(torch.bfloat16)


def _shard_worker(rank, world, path, ckdir):
    from lr2ppo_b200 import checkpoint
    from lr2ppo_b200.dist import GradSync
    dist.init_process_group("gloo", init_method=f"file://{path}", rank=rank, world_size=world)
    torch.manual_seed(7 + rank)
    net = _Net2()
    sync = GradSync(world)
    sync.broadcast_params(net)
    opt = _RowAdam(net.parameters())
    sync.attach(net, opt, shard_fc1=True)
    w = net.out_layer.fc1.weight
    per = w.shape[0] // world
    assert net._engine.fc1_rows == (rank * per, (rank + 1) * per) and opt.windows[id(w)] == (rank, world)
    assert sync.row_shards(net) == {"out_layer.fc1.weight": (rank * per, (rank + 1) * per)}
    w0 = w.detach().clone()
    shadow = net._engine.bank.get(w)
    torch.manual_seed(99)                                         # the global-batch gradients: same on every rank
    grads = {id(p): torch.randn_like(p) * world for p in net.parameters()}       # (grad_scale = 1/world undoes it)
    opt.step(grads, {id(w): shadow})
    sync.after_step(net)()                                        # in-place all-gather of the updated bf16 rows
    full = w0 - 0.1 * grads[id(w)] / world
    assert torch.equal(shadow, full.to(torch.bfloat16))           # every rank's GEMMs see the complete new weight
    own = slice(rank * per, (rank + 1) * per)
    stale = torch.ones(w.shape[0], dtype=torch.bool); stale[own] = False
    assert torch.equal(w.detach()[own], full[own]) and torch.equal(w.detach()[stale], w0[stale])   # master: own rows only
    # the foreign rows are stale now: a plain state_dict() must refuse instead of saving torn weights / moments
    for obj in (net, opt):
        try:
            obj.state_dict()
            raise AssertionError("state_dict() on stale row shards did not raise")
        except RuntimeError as e:
            assert "consolidate" in str(e)
    # sharded checkpoint: no consolidate; every rank writes its rows, rank 0 the rest
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: 1.0)
    checkpoint.save_sharded(ckdir, {"actor": net}, {"actor": opt}, {"actor": sched}, step=1, rank=rank, world=world,
                            row_shards={"actor": sync.row_shards(net)},
                            checkpointer=checkpoint.AsyncCheckpointer()).wait()
    dist.barrier()
    net2 = _Net2(); opt2 = _RowAdam(net2.parameters())
    sched2 = torch.optim.lr_scheduler.LambdaLR(opt2, lambda s: 1.0)
    assert checkpoint.load_sharded(ckdir, {"actor": net2}, {"actor": opt2}, {"actor": sched2})[0] == 1
    assert torch.equal(net2.out_layer.fc1.weight.detach(), full)  # complete fp32 master on every rank
    assert torch.equal(opt2.state[net2.out_layer.fc1.weight]["exp_avg"], 0.1 * grads[id(w)] / world)
    assert torch.equal(net2.other.weight.detach(), net.other.weight.detach())
    # consolidate(): the in-memory alternative (all-gathers master + moments in place)
    sync.consolidate(net, opt)
    assert torch.equal(w.detach(), full)
    assert torch.equal(opt.state[w]["exp_avg"], 0.1 * grads[id(w)] / world)
    assert torch.equal(net.state_dict()["out_layer.fc1.weight"], full) and opt.state_dict()["state"]   # allowed again
    dist.barrier()
    dist.destroy_process_group()


def test_row_sharded_optimizer_and_sharded_checkpoint_world2_gloo():
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_shard_worker, args=(2, os.path.join(d, "rdzv"), os.path.join(d, "ck")), nprocs=2, join=True)


# ---- K-split out_layer.fc1 (dist.Fc1Parallel): the exchange algebra on CPU, world 2 and 4, gloo -----------------
def _tp_worker(rank, world, path):
    from lr2ppo_b200.dist import Fc1Parallel
    dist.init_process_group("gloo", init_method=f"file://{path}", rank=rank, world_size=world)
    items, kb, hid = 3, 5, 7
    K1 = kb * world
    g = torch.Generator().manual_seed(1)                         # the same on every rank
    W = torch.randn(hid, K1, generator=g, dtype=torch.float64)
    X = [torch.randn(items, K1, generator=g, dtype=torch.float64) for _ in range(world)]
    dY = [torch.randn(items, hid, generator=g, dtype=torch.float64) for _ in range(world)]
    tp = Fc1Parallel(world, rank, (rank * kb, (rank + 1) * kb))
    k0, k1 = tp.cols
    # forward: my column block of everybody's rows, partial products, sum + scatter
    x_k = tp.scatter_cols(X[rank])
    assert torch.equal(x_k, torch.cat([x[:, k0:k1] for x in X]))
    y = tp.reduce_scatter(x_k @ W[:, k0:k1].t())
    assert torch.allclose(y, X[rank] @ W.t())
    # backward: dX of my column block for everybody's rows (complete sums), returned to the owners
    dy_all = tp.all_gather(dY[rank])
    assert torch.equal(dy_all, torch.cat(dY))
    dx = tp.gather_cols_async(dy_all @ W[:, k0:k1])()
    assert torch.allclose(dx, dY[rank] @ W)
    # weight gradient of the owned column block = global-batch gradient restricted to it
    g_own = dy_all.t() @ x_k
    g_ref = sum(dY[q].t() @ X[q] for q in range(world))
    assert torch.allclose(g_own, g_ref[:, k0:k1])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_fc1_k_split_exchange_algebra_gloo(world, tmp_path):
    mp.spawn(_tp_worker, args=(world, str(tmp_path / "pg")), nprocs=world, join=True)
