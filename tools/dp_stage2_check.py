"""2-GPU cross-check of data-parallel stage 2 (BASELINE configs[2]; torchrun --nproc-per-node 2 tools/dp_stage2_check.py):
one stages.reward_train_model step with dist.GradSync on two ranks (half of the pairs each) must give the update of
ONE rank stepping on the whole batch: same loss (mean of the two rank means), Adam first moments equal within the
bf16 tolerance (tests/parity.py: 4e-2 of each tensor's scale, 0.12 on the pair-cancellation tensors; tile shapes differ with the row count).  Dropout off, lr = 1e-3 constant.
Two steps are taken in both runs: the second step's loss only agrees if the forward weights followed the first update."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from lr2ppo_b200 import models, optim, stages
from lr2ppo_b200.dist import GradSync

PAIRS = 16            # global batch (pairs of tag orderings); 8 per rank


def build(dev):
    torch.manual_seed(7)
    margs = argparse.Namespace(mode="reg", labels_num=3, seq_length=196, max_imgs=16, visual_feat_dim=768)
    with torch.device(dev):
        model = models.PairClassifier(margs, margs)
    g = torch.Generator(device=dev).manual_seed(7)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "gamma" not in n and "beta" not in n:
                p.copy_(torch.randn(p.shape, generator=g, device=dev) * 0.02)
        for mod in model.modules():
            if isinstance(mod, torch.nn.LayerNorm):
                mod.weight.fill_(1.0)
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
    model.train()
    named = list(model.named_parameters())
    no_decay = ["bias", "gamma", "beta"]
    groups = [{"params": [p for n, p in named if not any(nd in n for nd in no_decay)], "weight_decay": 0.01},
              {"params": [p for n, p in named if any(nd in n for nd in no_decay)], "weight_decay": 0.0}]
    opt = optim.AdamW(groups, lr=1e-3, correct_bias=False)
    return model, opt, optim.get_constant_schedule(opt)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    g = torch.Generator().manual_seed(5)
    text = torch.randn(PAIRS, 2, 196, 768, generator=g)
    img = torch.randn(PAIRS, 1, 16, 768, generator=g).repeat(1, 2, 1, 1)
    tgts = torch.randint(0, 3, (PAIRS, 2), generator=g)
    flip = torch.rand(PAIRS, generator=g) < 0.5
    chosen = torch.where(flip[:, None], torch.tensor([[0, 1, 0, 1]]), torch.tensor([[1, 0, 0, 1]]))
    reject = torch.where(flip[:, None], torch.tensor([[0, 1, 1, 0]]), torch.tensor([[1, 0, 1, 0]]))
    args = argparse.Namespace(mode="reg")
    per = PAIRS // world
    sl = slice(rank * per, (rank + 1) * per)
    # data parallel: each rank steps on its slice
    model, opt, sch = build(dev)
    sync = GradSync(world)
    sync.broadcast_params(model)
    sync.attach(model, opt)
    for _ in range(2):      # the SECOND step's loss is computed with the weights the first step produced
        loss_dp, _ = stages.reward_train_model(args, model, opt, sch, text[sl].to(dev), img[sl].to(dev),
                                               tgts[sl].to(dev), chosen[sl].to(dev), reject[sl].to(dev),
                                               grad_sync=sync)
    loss_mean = loss_dp.detach().clone()
    dist.all_reduce(loss_mean)
    loss_mean /= world
    m_dp = {n: opt.state[p]["exp_avg"].clone() for n, p in model.named_parameters()}
    del model, opt, sch, sync
    torch.cuda.empty_cache()
    # reference: one rank, whole batch, no synchronisation
    model, opt, sch = build(dev)
    first = None
    for _ in range(2):
        loss_1, _ = stages.reward_train_model(args, model, opt, sch, text.to(dev), img.to(dev), tgts.to(dev),
                                              chosen.to(dev), reject.to(dev))
        first = loss_1.item() if first is None else first
    assert abs(loss_1.item() - first) > 1e-4 * abs(first), "the second step did not see the updated weights"
    bad = []
    top = max(opt.state[p]["exp_avg"].abs().max().item() for p in model.parameters())
    for n, p in model.named_parameters():
        a, b = m_dp[n].float(), opt.state[p]["exp_avg"].float()
        scale = b.abs().max().item()
        if scale < 1e-4 * top:
            continue          # mathematically-zero gradients (keys.bias: softmax is shift-invariant): noise on both sides
        err = (a - b).abs().max().item() / scale
        # the bf16 bounds of tests/parity.py: 4e-2 per gradient tensor, 0.12 where the chosen / reject passes cancel
        tol = 0.12 if any(k in n for k in ("pos_emb", "out_layer.fc2.bias", "xitt.")) else 4e-2
        if err > tol:
            bad.append((n, err))
    lerr = abs(loss_mean.item() - loss_1.item()) / max(1e-6, abs(loss_1.item()))
    print(f"[rank {rank}] loss dp-mean {loss_mean.item():.6f} vs single {loss_1.item():.6f} (rel {lerr:.2e}); "
          f"moment mismatches: {bad[:5]}", flush=True)
    ok = torch.tensor([0 if (bad or lerr > 2e-2) else 1], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DP_STAGE2_CHECK", "PASS" if ok.item() == 1 else "FAIL", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
