"""Asynchronous checkpoint + bit-exact resume on the device path (FusedAdamW moments, bf16 shadows, schedule)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from lr2ppo_b200 import checkpoint
from lr2ppo_b200.optim import FusedAdamW, get_linear_schedule_with_warmup


class _Net(torch.nn.Module):
    def __init__(self):
        super().__init__()
        g = torch.Generator().manual_seed(4)
        self.w = torch.nn.Parameter(torch.randn(300, 4099, generator=g) * 0.02)
        self.b = torch.nn.Parameter(torch.randn(4099, generator=g) * 0.02)


def _make():
    net = _Net().cuda()
    opt = FusedAdamW([{"params": [net.w], "weight_decay": 0.01}, {"params": [net.b], "weight_decay": 0.0}], lr=1e-3,
                     correct_bias=False, shadow_bf16=True)
    sch = get_linear_schedule_with_warmup(opt, 2, 20)
    return net, opt, sch


def _steps(net, opt, sch, first, n):
    for s in range(first, first + n):
        g = torch.Generator(device="cuda").manual_seed(100 + s)
        net.w.grad = torch.randn(net.w.shape, generator=g, device="cuda") * 0.01
        net.b.grad = torch.randn(net.b.shape, generator=g, device="cuda") * 0.01
        opt.step(); sch.step()


def test_async_save_and_bit_exact_resume(tmp_path):
    net, opt, sch = _make()
    _steps(net, opt, sch, 0, 3)
    path = str(tmp_path / "resume.pt")
    ck = checkpoint.save_training_state(path, {"net": net}, {"opt": opt}, {"sch": sch}, step=3)
    _steps(net, opt, sch, 3, 3)                # keeps training while the file is being written
    ck.wait()
    net2, opt2, sch2 = _make()
    step, _ = checkpoint.load_training_state(path, {"net": net2}, {"opt": opt2}, {"sch": sch2})
    assert step == 3
    _steps(net2, opt2, sch2, 3, 3)
    assert torch.equal(net.w.detach(), net2.w.detach()) and torch.equal(net.b.detach(), net2.b.detach())
    assert torch.equal(opt.state_for(net.w)["exp_avg_sq"], opt2.state_for(net2.w)["exp_avg_sq"])
    assert torch.equal(opt.shadow_of(net.w), opt2.shadow_of(net2.w))
    # weights-only export in the reference's format
    out = str(tmp_path / "finetuned_model.bin")
    checkpoint.save_model(net, out).wait()
    sd = torch.load(out, map_location="cpu")
    assert set(sd) == {"w", "b"} and sd["w"].dtype == torch.float32 and torch.equal(sd["w"], net.w.detach().cpu())
