"""What every stage script's main() does before it builds its model (finetune/ppo.py:702-763)."""
import torch

from .. import cli, runtime
from ..tokenizers import str2tokenizer


def prologue(stage, argv=None):
    """-> (args, vit_args, world, rank): flags parsed, JSON configs merged (defaults < file < command line), process
    group up, RNGs seeded with seed + rank, tokenizers constructed (their result is not used afterwards: the text and
    image towers' outputs are pre-extracted features)."""
    import sys
    args = cli.stage_parser(stage).parse_args(argv)
    vit_args = cli.vit_namespace(args)
    given = sys.argv if argv is None else list(argv)
    args = cli.load_hyperparam(args, given)
    vit_args = cli.load_hyperparam(vit_args, given)
    args.labels_num = 3
    runtime.init_distributed_mode(args)
    runtime.setup_seed(args.seed + runtime.get_rank())
    args.is_master = runtime.is_main_process()
    args.tokenizer = str2tokenizer[args.tokenizer](args)
    vit_args.tokenizer = str2tokenizer[vit_args.tokenizer](vit_args)
    if not torch.cuda.is_available():
        raise RuntimeError("the LR2PPO stage scripts run on CUDA (sm_100a) only; there is no CPU fallback")
    args.device = torch.device("cuda", args.gpu)
    return args, vit_args, runtime.get_world_size(), runtime.get_rank()


def to_device(args, *tensors):
    """Asynchronous upload of one (pinned) loader batch."""
    return tuple(t.to(args.device, non_blocking=True) for t in tensors)


def grad_sync_for(world):
    """north_star's data-parallel gradient averaging (LR2_GRAD_SYNC=1).  Off by default: the reference trains
    independent replicas in its multimodal scripts (no DDP wrap) and parity is defined on that."""
    import os
    if world > 1 and os.environ.get("LR2_GRAD_SYNC", "0") == "1":
        from ..dist import GradSync
        return GradSync(world)
    return None


def replicated(sync, *modules):
    """Context for code the ranks do not run in lock-step (evaluation): K-split modules temporarily run replicated."""
    import contextlib
    return sync.replicated(*modules) if sync is not None else contextlib.nullcontext()


def save_if_best(args, sync, pairs, improved, model, path):
    """Rank 0 decides (`improved`, evaluated on rank 0 only like the reference's `if args.is_master and result >
    best_result`); with sharded data-parallel optimizers every rank must take part in completing the fp32 masters
    (dist.GradSync.consolidate) before rank 0 writes the file, so the decision is broadcast first."""
    import torch
    import torch.distributed as dist
    from .. import checkpoint
    if sync is not None:
        flag = torch.tensor([1 if (args.is_master and improved) else 0], device=args.device)
        dist.broadcast(flag, 0)
        improved = bool(flag.item())
        if improved:
            for module, optimizer in pairs:
                sync.consolidate(module, optimizer)
    if args.is_master and improved:
        checkpoint.save_model(model, path)
    return improved
