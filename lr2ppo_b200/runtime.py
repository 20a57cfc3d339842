"""Process-level plumbing of the stage scripts: distributed start-up and rank helpers (finetune/misc.py:22-107),
seeding (tencentpretrain/utils/seed.py), the logger (tencentpretrain/utils/logging.py:4-19) and parameter
initialisation / checkpoint loading (finetune/ppo.py:358-375, finetune/pointwise.py:239-272).

One process per GPU under torchrun; NCCL over NVLink for the collectives.  `LR2_DIST_BACKEND=gloo` lets the same
entry points start on a CPU-only host for argument / data-path tests (the models themselves need a B200)."""
import builtins
import datetime
import logging
import os
import random

import numpy as np
import torch
import torch.distributed as dist


def is_dist_avail_and_initialized():
    return dist.is_available() and dist.is_initialized()


def get_world_size():
    return dist.get_world_size() if is_dist_avail_and_initialized() else 1


def get_rank():
    return dist.get_rank() if is_dist_avail_and_initialized() else 0


def is_main_process():
    return get_rank() == 0


def mkdir(path):
    os.makedirs(path, exist_ok=True)


def setup_seed(seed):
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    np.random.seed(seed)
    random.seed(seed)


def set_seed(seed=7):
    os.environ["PYTHONHASHSEED"] = str(seed)
    setup_seed(seed)


def setup_for_distributed(is_master):
    """print() only on the master rank, prefixed with the wall-clock time (`force=True` overrides)."""
    plain = builtins.print

    def rank0_print(*args, **kwargs):
        force = kwargs.pop("force", False) or get_world_size() > 8
        if is_master or force:
            plain("[{}] ".format(datetime.datetime.now().time()), end="")
            plain(*args, **kwargs)

    builtins.print = rank0_print


def init_distributed_mode(args):
    """env:// rendezvous from torchrun's RANK / WORLD_SIZE / LOCAL_RANK, NCCL backend, barrier, rank-0 printing."""
    args.rank = int(os.environ["RANK"])
    args.world_size = int(os.environ["WORLD_SIZE"])
    args.gpu = int(os.environ["LOCAL_RANK"])
    args.distributed = True
    args.dist_backend = os.environ.get("LR2_DIST_BACKEND", "nccl")
    kw = {}
    if args.dist_backend == "nccl":
        torch.cuda.set_device(args.gpu)
        kw["device_id"] = torch.device("cuda", args.gpu)
    print("| distributed init (rank {}): {}, gpu {}".format(args.rank, args.dist_url, args.gpu), flush=True)
    dist.init_process_group(backend=args.dist_backend, init_method=args.dist_url, world_size=args.world_size,
                            rank=args.rank, **kw)
    dist.barrier()
    setup_for_distributed(args.rank == 0)


def init_logger(args):
    """Root logger: console + optional --log_path file, `[time LEVEL] message` lines (diff-able with logs/)."""
    fmt = logging.Formatter("[%(asctime)s %(levelname)s] %(message)s")
    logger = logging.getLogger()
    logger.setLevel(args.log_level)
    console = logging.StreamHandler()
    console.setFormatter(fmt)
    logger.handlers = [console]
    if args.log_path is not None:
        fh = logging.FileHandler(args.log_path, encoding="UTF-8")
        fh.setLevel(args.log_file_level)
        fh.setFormatter(fmt)
        logger.addHandler(fh)
    return logger


def init_normal_(model):
    """N(0, 0.02) for everything but gamma / beta (finetune/ppo.py:363-365)."""
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "gamma" not in n and "beta" not in n:
                p.normal_(0, 0.02)


def load_strict_or_init(path, model):
    """Stage 3 / eval: strict load of a checkpoint written by save_model, else random init
    (finetune/ppo.py:358-375)."""
    if path is not None:
        model.load_state_dict(torch.load(path, map_location="cpu"), strict=True)
    else:
        init_normal_(model)


def load_towers_or_init(args, model):
    """Stage 1 / 2: the --pretrained_model_path / --vit_pretrained_model_path files hold TOWER weights; the fusion model
    has no matching keys, so the reference's strict=False loads change nothing (SURVEY.md §0 fact 2,
    finetune/pointwise.py:239-266) and, having taken that branch, it skips the random initialisation too: the model
    keeps torch's default nn.Linear / nn.LayerNorm init.  Reproduced: keys are matched by name, unmatched ones
    reported, and nothing is re-initialised when a path is given."""
    if args.pretrained_model_path is None:
        init_normal_(model)
        return
    own = model.state_dict()
    for tag, path, prefix in (("text", args.pretrained_model_path, ""),
                              ("img", getattr(args, "vit_pretrained_model_path", None), "vit_")):
        if path is None or not os.path.exists(path):
            print(f"{tag} tower checkpoint {path!r} not found: nothing to load (the fusion model holds no tower keys)")
            continue
        ckpt = {prefix + k: v for k, v in torch.load(path, map_location="cpu").items()}
        hit = {k: v for k, v in ckpt.items() if k in own and own[k].shape == v.shape}
        model.load_state_dict(hit, strict=False)
        print(f"{tag} tower checkpoint {path}: {len(hit)} of {len(ckpt)} tensors matched the fusion model")
