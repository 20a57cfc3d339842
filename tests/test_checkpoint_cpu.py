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
