"""checkpoint.py host logic: the asynchronous saver writes the reference's format atomically and a
training state round-trips.  Runs without a GPU (synchronous copies); the GPU variant is in test_checkpoint_gpu.py."""
import os

import torch

from lr2ppo_b200 import checkpoint


def _model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.GELU(), torch.nn.Linear(5, 3))


def test_save_model_matches_reference_format(tmp_path):
    m = _model()
    path = str(tmp_path / "finetuned_model.bin")
    ck = checkpoint.save_model(m, path)
    with torch.no_grad():
        for p in m.parameters():          # later modification must not leak into the snapshot
            p.add_(1.0)
    ck.wait()
    assert not [f for f in os.listdir(tmp_path) if ".tmp." in f]
    ref = _model()
    sd = torch.load(path, map_location="cpu")            # exactly how finetune/ppo.py:360-361 reads it
    ref2 = _model()
    ref2.load_state_dict(sd, strict=True)
    for a, b in zip(ref.parameters(), ref2.parameters()):
        assert torch.equal(a, b)
    # buffers are reused: a second save of different values overwrites the file atomically
    checkpoint.save_model(m, path).wait()
    sd2 = torch.load(path, map_location="cpu")
    assert torch.equal(sd2["0.weight"], m[0].weight.detach())


def test_training_state_round_trip(tmp_path):
    m = _model()
    opt = torch.optim.AdamW(m.parameters(), lr=1e-2)
    sch = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: 1.0 / (1 + s))
    x = torch.randn(4, 7)

    def step():
        opt.zero_grad()
        (m(x) ** 2).sum().backward()
        opt.step(); sch.step()

    step(); step()
    path = str(tmp_path / "state.pt")
    checkpoint.save_training_state(path, {"m": m}, {"o": opt}, {"s": sch}, step=2, extra={"best": 0.5}).wait()
    step(); step()
    want = [p.detach().clone() for p in m.parameters()]
    m2 = _model()
    opt2 = torch.optim.AdamW(m2.parameters(), lr=1e-2)
    sch2 = torch.optim.lr_scheduler.LambdaLR(opt2, lambda s: 1.0 / (1 + s))
    k, extra = checkpoint.load_training_state(path, {"m": m2}, {"o": opt2}, {"s": sch2})
    assert k == 2 and extra == {"best": 0.5}
    m, opt, sch = m2, opt2, sch2
    step(); step()
    for a, b in zip(want, m2.parameters()):
        assert torch.equal(a, b)


# ---- sharded training state (row-sharded out_layer.fc1 optimizer) ----------------------------------------------
class _Fusion(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.out_layer = torch.nn.Module()
        self.out_layer.fc1 = torch.nn.Linear(24, 8)
        self.head = torch.nn.Linear(8, 1)

    def forward(self, x):
        return self.head(torch.relu(self.out_layer.fc1(x)))


def _trained(seed, steps=3):
    torch.manual_seed(seed)
    m = _Fusion()
    opt = torch.optim.AdamW(m.parameters(), lr=1e-2)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: 1.0 / (1 + s))
    for _ in range(steps):
        opt.zero_grad()
        m(torch.randn(5, 24)).sum().backward()
        opt.step(); sched.step()
    return m, opt, sched


def _rank_view(truth, world, rank):
    """What rank `rank` holds in row-sharded mode: its own rows of fc1 (weight + moments) are authoritative, the
    other rows are stale (poisoned here so that a loader which used them would be caught)."""
    import copy
    m, opt, sched = truth
    m2 = copy.deepcopy(m)
    opt2 = torch.optim.AdamW(m2.parameters(), lr=1e-2)
    sched2 = torch.optim.lr_scheduler.LambdaLR(opt2, lambda s: 1.0 / (1 + s))    # (its constructor resets the lr)
    opt2.load_state_dict(copy.deepcopy(opt.state_dict()))
    sched2.load_state_dict(sched.state_dict())
    w = m2.out_layer.fc1.weight
    rows = w.shape[0] // world
    r0, r1 = rank * rows, (rank + 1) * rows
    stale = torch.ones(w.shape[0], dtype=torch.bool); stale[r0:r1] = False
    with torch.no_grad():
        w[stale] = float("nan")
        opt2.state[w]["exp_avg"][stale] = float("nan")
        opt2.state[w]["exp_avg_sq"][stale] = float("nan")
    return m2, opt2, sched2, {"out_layer.fc1.weight": (r0, r1)}


def test_sharded_save_load_roundtrip_and_export(tmp_path):
    from lr2ppo_b200 import checkpoint
    truth = _trained(0)
    world = 4
    d = str(tmp_path / "ck")
    for rank in range(world):
        m, opt, sched, rows = _rank_view(truth, world, rank)
        ck = checkpoint.save_sharded(d, {"actor": m}, {"actor": opt}, {"actor": sched}, step=7, rank=rank, world=world,
                                     row_shards={"actor": rows}, checkpointer=checkpoint.AsyncCheckpointer(),
                                     extra={"best": 0.5})
        ck.wait()
    m3, opt3, sched3 = _trained(1, steps=1)                       # different weights / state, to be overwritten
    step, extra = checkpoint.load_sharded(d, {"actor": m3}, {"actor": opt3}, {"actor": sched3})
    assert step == 7 and extra == {"best": 0.5}
    m, opt, sched = truth
    for (n, a), (_, b) in zip(m.state_dict().items(), m3.state_dict().items()):
        assert torch.equal(a, b), n                               # complete fp32 parameters, no stale row used
    for p, q in zip(m.parameters(), m3.parameters()):
        for k in ("exp_avg", "exp_avg_sq"):
            assert torch.equal(opt.state[p][k], opt3.state[q][k])
        assert float(opt.state[p]["step"]) == float(opt3.state[q]["step"])
    assert sched3.state_dict()["last_epoch"] == sched.state_dict()["last_epoch"]
    # one more identical step on both: the resumed run continues bit-identically
    x = torch.randn(5, 24)
    for mm, oo in ((m, opt), (m3, opt3)):
        oo.zero_grad(); mm(x).sum().backward(); oo.step()
    assert torch.equal(m.out_layer.fc1.weight, m3.out_layer.fc1.weight)
    # export: the reference's single-file format with the reference's key names
    out = str(tmp_path / "actor.bin")
    checkpoint.export_model(d, "actor", out)
    sd = torch.load(out)
    fresh = _Fusion()
    fresh.load_state_dict(sd, strict=True)
    assert set(sd) == {"out_layer.fc1.weight", "out_layer.fc1.bias", "head.weight", "head.bias"}
    assert all(v.dtype == torch.float32 for v in sd.values())


def test_sharded_load_detects_missing_or_mismatched_shards(tmp_path):
    import os
    import pytest
    from lr2ppo_b200 import checkpoint
    truth = _trained(0)
    d = str(tmp_path / "ck")
    for rank in range(2):
        m, opt, sched, rows = _rank_view(truth, 2, rank)
        checkpoint.save_sharded(d, {"a": m}, {"a": opt}, {"a": sched}, step=3, rank=rank, world=2,
                                row_shards={"a": rows}, checkpointer=checkpoint.AsyncCheckpointer()).wait()
    m3, opt3, sched3 = _trained(1, steps=1)
    os.rename(os.path.join(d, "shard-00001-of-00002.pt"), os.path.join(d, "hidden"))
    with pytest.raises(FileNotFoundError):
        checkpoint.load_sharded(d, {"a": m3}, {"a": opt3}, {"a": sched3})
    # a shard from another step must not be mixed in
    m, opt, sched, rows = _rank_view(truth, 2, 1)
    checkpoint.save_sharded(d, {"a": m}, {"a": opt}, {"a": sched}, step=4, rank=1, world=2, row_shards={"a": rows},
                            checkpointer=checkpoint.AsyncCheckpointer()).wait()
    with pytest.raises(RuntimeError):
        checkpoint.load_sharded(d, {"a": m3}, {"a": opt3}, {"a": sched3})


def test_replicated_model_through_sharded_api(tmp_path):
    """row_shards = {}: rank 0 writes everything, the other ranks write empty shards; same loader."""
    from lr2ppo_b200 import checkpoint
    m, opt, sched = _trained(2)
    d = str(tmp_path / "ck")
    for rank in range(2):
        checkpoint.save_sharded(d, {"m": m}, {"m": opt}, {"m": sched}, step=1, rank=rank, world=2, row_shards={},
                                checkpointer=checkpoint.AsyncCheckpointer()).wait()
    m3, opt3, sched3 = _trained(3, steps=1)
    assert checkpoint.load_sharded(d, {"m": m3}, {"m": opt3}, {"m": sched3})[0] == 1
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), m3.state_dict().values()))


def test_sharded_checkpoint_with_column_blocks(tmp_path):
    """K-split data parallel (dist.Fc1Parallel): every rank owns a COLUMN block of out_layer.fc1 -- descriptors
    (dim, lo, hi) -- and the loader must reassemble complete tensors from them."""
    import torch
    from lr2ppo_b200 import checkpoint

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.fc1 = torch.nn.Linear(12, 5)
            self.head = torch.nn.Linear(5, 1)

    torch.manual_seed(0)
    truth = Net()
    world, kb = 3, 4
    full_m = torch.randn(5, 12)
    d = str(tmp_path / "ck")
    for rank in range(world):
        net = Net()
        net.load_state_dict(truth.state_dict())
        opt = torch.optim.SGD(net.parameters(), lr=0.1, momentum=0.9)
        for p in net.parameters():
            opt.state[p] = {"momentum_buffer": torch.zeros_like(p), "exp_avg": torch.zeros_like(p),
                            "exp_avg_sq": torch.zeros_like(p)}
        lo, hi = rank * kb, (rank + 1) * kb
        with torch.no_grad():                                    # foreign column blocks are stale (poisoned)
            mask = torch.ones(12, dtype=torch.bool); mask[lo:hi] = False
            net.fc1.weight[:, mask] = float("nan")
            opt.state[net.fc1.weight]["exp_avg"][:, lo:hi] = full_m[:, lo:hi]
            opt.state[net.fc1.weight]["exp_avg"][:, mask] = float("nan")
        sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: 1.0)
        checkpoint.save_sharded(d, {"m": net}, {"m": opt}, {"m": sched}, step=3, rank=rank, world=world,
                                row_shards={"m": {"fc1.weight": (1, lo, hi)}},
                                checkpointer=checkpoint.AsyncCheckpointer()).wait()
    net2 = Net()
    opt2 = torch.optim.SGD(net2.parameters(), lr=0.1, momentum=0.9)
    for p in net2.parameters():
        opt2.state[p] = {"momentum_buffer": torch.zeros_like(p), "exp_avg": torch.zeros_like(p),
                         "exp_avg_sq": torch.zeros_like(p)}
    sched2 = torch.optim.lr_scheduler.LambdaLR(opt2, lambda s: 1.0)
    assert checkpoint.load_sharded(d, {"m": net2}, {"m": opt2}, {"m": sched2})[0] == 3
    assert torch.equal(net2.fc1.weight.detach(), truth.fc1.weight.detach())
    assert torch.equal(opt2.state[net2.fc1.weight]["exp_avg"], full_m)
    sd = checkpoint.export_model(d, "m", str(tmp_path / "export.bin"))
    assert torch.equal(sd["fc1.weight"], truth.fc1.weight.detach())
