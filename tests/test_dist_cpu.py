"""N>1 host logic on CPU (gloo, world_size 2): the data-parallel gradient synchronisation of lr2ppo_b200.dist.
The CUDA kernels are not involved; this checks the collective plumbing and the algebra that lets out_layer.fc1
all-gather its wgrad operands instead of all-reducing the 2 GB gradient."""
import os
import tempfile

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class _Eng:
    def __init__(self, m):
        self.m, self.dp_gather, self.bank = m, None, type("B", (), {})()


class _Net(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.out_layer = torch.nn.Module()
        self.out_layer.fc1 = torch.nn.Linear(24, 8)
        self.other = torch.nn.Linear(8, 4)
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
    dist.barrier()
    dist.destroy_process_group()


def test_gradsync_world2_gloo():
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(2, os.path.join(d, "rdzv")), nprocs=2, join=True)
