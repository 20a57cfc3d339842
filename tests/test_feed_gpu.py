"""DeviceFeeder / StatsReader (lr2ppo_b200/feed.py): double-buffered pinned H2D feed and lagged D2H read-back."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from lr2ppo_b200.feed import DeviceFeeder, StatsReader


def test_feeder_delivers_batches_in_order_while_compute_runs():
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(0)
    pool = [(torch.randn(24, 2, 196, 768, generator=g).pin_memory(),
             torch.randint(0, 3, (24, 2), generator=g).pin_memory()) for _ in range(5)]
    feeder = DeviceFeeder(pool[0], dev)
    dst = tuple(torch.empty_like(t, device=dev) for t in pool[0])
    busy = torch.randn(4096, 4096, device=dev)
    feeder.stage(pool[0])
    sums = []
    for i in range(12):
        feeder.next_into(dst)
        feeder.stage(pool[(i + 1) % len(pool)])
        busy = busy @ busy * 1e-4                      # keep the compute stream busy while the next copy runs
        sums.append((dst[0].double().sum(), dst[1].sum()))
    torch.cuda.synchronize()
    for i, (a, b) in enumerate(sums):
        ref = pool[i % len(pool)]
        assert abs(a.item() - ref[0].double().sum().item()) < 1e-6
        assert b.item() == ref[1].sum().item()


def test_feeder_refuses_unpinned_and_overflow():
    dev = torch.device("cuda", 0)
    ex = (torch.zeros(8).pin_memory(),)
    f = DeviceFeeder(ex, dev)
    with pytest.raises(RuntimeError):
        f.stage((torch.zeros(8),))
    f = DeviceFeeder(ex, dev)
    f.stage(ex); f.stage(ex)
    with pytest.raises(RuntimeError):
        f.stage(ex)
    with pytest.raises(RuntimeError):
        DeviceFeeder(ex, dev).next_into((torch.zeros(8, device=dev),))


def test_stats_reader_lags_one_step_and_flushes():
    r = StatsReader(4)
    outs = []
    for i in range(6):
        prev = r.push(torch.full((4,), float(i), device="cuda"))
        outs.append(None if prev is None else prev.clone())
    assert outs[0] is None
    for i in range(1, 6):
        assert torch.equal(outs[i], torch.full((4,), float(i - 1)))
    last = r.flush()
    assert len(last) == 1 and torch.equal(last[0], torch.full((4,), 5.0))
