"""Evaluation -- `sh ppo_eval.sh NAME` (finetune/ppo_eval.py:492-587): strict load of the stage-3 checkpoint
(`actor.*` + `critic.*` keys) into ActorCritic, NDCG@{1,3,5,10,20,all} over the validation clips, and the per-clip
dump `case/ppo_cases.json`."""
import torch

from .. import data, ppo, runtime
from . import common


def main(argv=None):
    args, vit_args, num_tasks, global_rank = common.prologue("ppo_eval", argv)
    model = ppo.ActorCritic(args, vit_args)
    runtime.load_strict_or_init(args.pretrained_model_path, model)      # the full ActorCritic state_dict, strict
    if args.is_master:
        args.logger = runtime.init_logger(args)
    model = model.to(args.device)
    valset = data.EvalClips(args, args.dev_path, is_train=False)
    val_loader = data.get_dataloader(args, valset, num_tasks, global_rank, is_train=False)
    args.model = model
    with torch.no_grad():
        result = ppo.evaluate(args, val_loader, 0, split="val", num_tasks=num_tasks, cases_path="case/ppo_cases.json")
    return result


if __name__ == "__main__":
    main()
