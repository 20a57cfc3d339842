"""Multi-GPU cross-check (torchrun --nproc-per-node 2|4|8) of the three data-parallel modes of dist.GradSync on
identical models / batches, dropout off, constant lr = 2e-5, three eager stage-3 steps each:

  replicated   every rank updates the whole out_layer.fc1 from all-gathered wgrad operands           (reference)
  gather       row-sharded optimizer + all-gather of the updated bf16 rows (round 1)                  == replicated, bit for bit
  tp           column-sharded optimizer + K-split fc1 (dist.Fc1Parallel, the default)                ~= replicated

`tp` changes the arithmetic in one place: the pre-activation of fc1 is an fp32 sum over the ranks of per-rank partial
products (reduce-scatter) instead of one split-K GEMM -- a different summation order, then the same bf16 rounding --
so it is held to the bf16 tolerance instead of bit-equality: statistics of the last step 2e-2, every Adam first moment 4e-2 of its scale (tests/parity.py), and --
because forward weights that silently stopped following the optimizer would pass a moments-only check -- the bf16
weights every rank actually multiplies with must equal the rounded fp32 masters after consolidation.

Both sharded modes also write a per-rank sharded checkpoint (checkpoint.save_sharded with GradSync.row_shards: row
blocks in gather mode, column blocks in tp mode) before consolidating; rank 0 reassembles the files and compares
them bit for bit with the consolidated master weights and Adam moments."""
import argparse
import os
import shutil
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from lr2ppo_b200 import checkpoint, ppo
from lr2ppo_b200.dist import GradSync


def build(seed, dev):
    torch.manual_seed(seed)
    margs = argparse.Namespace(mode="reg", labels_num=3, seq_length=196, max_imgs=16, visual_feat_dim=768)
    with torch.device(dev):
        model = ppo.ActorCritic(margs, margs)
        reward = ppo.Reward(margs, margs)
    g = torch.Generator(device=dev).manual_seed(seed)
    with torch.no_grad():
        for m in (model, reward):
            for n, p in m.named_parameters():
                if "gamma" not in n and "beta" not in n:
                    p.copy_(torch.randn(p.shape, generator=g, device=dev) * 0.02)
            for mod in m.modules():
                if isinstance(mod, torch.nn.LayerNorm):
                    mod.weight.fill_(1.0)
                if isinstance(mod, torch.nn.Dropout):
                    mod.p = 0.0
    model.eval(); reward.eval()
    return model, reward


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    # lr = 2e-5: large enough that every weight visibly moves in three steps, small enough that the sign-like first
    # Adam steps do not amplify rounding differences between the modes into different trajectories
    hp = argparse.Namespace(learning_rate=2e-5, critic_learning_rate=2e-5, optimizer="adamw", scheduler="constant",
                            train_steps=100, warmup=0.1, kl_div_loss_weight=0.001, entropy_weight=0.001,
                            value_clip=0.5, mode="reg", fc1_grad_bf16=True)
    g = torch.Generator().manual_seed(100 + rank)
    batches = [(torch.randn(24, 2, 196, 768, generator=g).to(dev),
                torch.randn(24, 1, 16, 768, generator=g).repeat(1, 2, 1, 1).to(dev),
                torch.randint(0, 3, (24, 2), generator=g).to(dev)) for _ in range(3)]
    results, ck_bad = {}, []
    ck_root = os.environ.get("LR2_CHECK_DIR", "/tmp/lr2_dp_shard_check")
    if rank == 0:
        shutil.rmtree(ck_root, ignore_errors=True)
    dist.barrier()
    for mode in ("replicated", "gather", "tp"):
        model, reward = build(7, dev)
        opt, copt, sch, csch = ppo.build_optimizer(hp, model)
        sync = GradSync(world)
        sync.broadcast_params(model); sync.broadcast_params(reward)
        sync.attach(model.actor, opt, shard_fc1=mode != "replicated", tensor_parallel=mode == "tp")
        sync.attach(model.critic, copt, shard_fc1=mode != "replicated", tensor_parallel=mode == "tp")
        assert (model.actor._engine.fc1_rows is not None) == (mode == "gather")
        assert (model.actor._engine.tp is not None) == (mode == "tp")
        stats = None
        for text, img, tgts in batches:
            mem = ppo.rollout(model, reward, text, img, tgts)
            model.train()
            stats = ppo.update_batch(hp, model, opt, copt, mem, sync)
            model.eval()
        torch.cuda.synchronize()
        ck_dir = None
        if mode != "replicated":
            # sharded checkpoint on real ranks: every rank writes its own block of out_layer.fc1 (rows in gather mode,
            # columns in tp mode) BEFORE any consolidation; rank 0 adds the replicated rest.  Read back below.
            ck_dir = os.path.join(ck_root, mode)
            # (the actor only: 6 GB of weights + moments per mode is enough disk traffic for a check)
            shards = {"actor": sync.row_shards(model.actor)}
            assert all(len(v) == 1 for v in shards.values()), shards
            checkpoint.save_sharded(ck_dir, {"actor": model.actor}, {"actor": opt}, {"actor": sch}, step=3, rank=rank,
                                    world=world, row_shards=shards).wait()
            dist.barrier()
            sync.consolidate(model.actor, opt); sync.consolidate(model.critic, copt)
        if mode == "tp":
            sync.gather_shadow(model.actor); sync.gather_shadow(model.critic)
        rec = {"stats": stats.float().cpu()}
        for tag, mod, o in (("actor", model.actor, opt), ("critic", model.critic, copt)):
            for n, p in mod.named_parameters():
                rec[f"{tag}.{n}"] = p.detach().clone() if p.numel() < 50_000_000 else None
                if p.numel() < 50_000_000:
                    rec[f"{tag}.{n}.m"] = o.state_for(p)["exp_avg"].clone()
            w = mod.out_layer.fc1.weight
            rec[f"{tag}.fc1.shadow"] = mod._engine.bank.get(w).clone()
            rec[f"{tag}.fc1.master"] = w.detach().clone()
            rec[f"{tag}.fc1.m"] = o.state_for(w)["exp_avg"].clone()
        if ck_dir is not None and rank == 0:
            # the shard files reassemble to exactly the consolidated tensors: master weights and both Adam moments
            common, parts = checkpoint._read_sharded(ck_dir, "cpu")
            for tag, mod, o in (("actor", model.actor, opt),):
                name = "out_layer.fc1.weight"
                assert list(common["sharded"][tag]) == [name] and name not in common["models"][tag]
                w = mod.out_layer.fc1.weight
                want = {"param": w.detach().cpu(), "exp_avg": o.state_for(w)["exp_avg"].cpu(),
                        "exp_avg_sq": o.state_for(w)["exp_avg_sq"].cpu()}
                for kind, ref in want.items():
                    got = checkpoint._assemble(list(w.shape), [(sh["rows"][tag][name], sh[kind][tag][name]) for sh in parts])
                    if not torch.equal(got, ref):
                        ck_bad.append((f"{mode}:{tag}.{kind} from shard files != consolidated", 1.0))
                dims = {checkpoint._block(sh["rows"][tag][name])[0] for sh in parts}
                assert dims == ({1} if mode == "tp" else {0}), dims            # tp: column blocks, gather: row blocks
            del common, parts
            shutil.rmtree(ck_dir, ignore_errors=True)
            print(f"[rank 0] {mode}: sharded checkpoint of {world} ranks reassembles to the consolidated fc1 weights "
                  f"and moments: {not ck_bad}", flush=True)
        results[mode] = rec
        del model, reward, opt, copt, sync
        torch.cuda.empty_cache()
    a, b, c = results["replicated"], results["gather"], results["tp"]
    bad = list(ck_bad)
    for k in a:
        if a[k] is None:
            continue
        if not torch.equal(a[k], b[k]):
            d = (a[k].float() - b[k].float()).abs().max().item()
            bad.append((k, d))
    moved = (a["actor.fc1.master"] - build(7, dev)[0].actor.out_layer.fc1.weight.detach()).abs().max().item()
    print(f"[rank {rank}] gather vs replicated: compared {len(a)} tensors, mismatches: {bad[:6]}; fc1 moved by "
          f"{moved:.3e}; stats {a['stats'][:3].tolist()} vs {b['stats'][:3].tolist()}", flush=True)
    # ---- tp vs replicated: bf16 tolerance
    worst = ("", 0.0)
    top = {tag: max(a[k].float().abs().max().item() for k in a if k.startswith(tag) and k.endswith(".m"))
           for tag in ("actor", "critic")}
    for k in a:
        if not k.endswith(".m") and k != "stats":
            continue
        ref, got = a[k].float(), c[k].float()
        scale = ref.abs().max().item()
        if k != "stats" and scale < 1e-4 * top[k.split(".")[0]]:
            continue       # mathematically-zero gradients (keys.bias, the actor's head.bias): fp32 noise on both sides
        err = (ref - got).abs().max().item() / max(scale, 1e-30)
        if err > worst[1]:
            worst = (k, err)
        tol = 2e-2 if k == "stats" else 4e-2
        if err > tol:
            bad.append(("tp:" + k, err))
    for tag in ("actor", "critic"):
        # the weights the GEMMs multiply with are the rounded masters (nothing went stale), in every mode
        for res, name in ((a, "replicated"), (b, "gather"), (c, "tp")):
            if not torch.equal(res[f"{tag}.fc1.shadow"], res[f"{tag}.fc1.master"].bfloat16()):
                bad.append((f"{name}:{tag}.fc1 shadow != bf16(master)", 1.0))
        same = (a[f"{tag}.fc1.master"] == c[f"{tag}.fc1.master"]).float().mean().item()
        print(f"[rank {rank}] tp {tag}: {same:.4f} of the fc1 master weights bit-equal to replicated after 3 steps",
              flush=True)
    print(f"[rank {rank}] tp vs replicated: worst moment / stat error {worst[1]:.4f} on {worst[0]}; stats "
          f"{a['stats'][:3].tolist()} vs {c['stats'][:3].tolist()}", flush=True)
    dist.barrier()
    ok = torch.tensor([0 if bad else 1], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DP_SHARD_CHECK", "PASS" if ok.item() == 1 else "FAIL", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
