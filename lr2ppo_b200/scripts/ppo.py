"""Stage 3 -- `sh ppo.sh NAME` (finetune/ppo.py:702-914).  Loop shape, logging lines and checkpoint format are the
reference's; the work inside runs on the B200 path:

  * batches are uploaded asynchronously from pinned memory, img_emb un-repeated;
  * the rollout of every batch and the update of every stored batch replay CUDA graphs (ppo.GraphedCycle), the
    stored batches live in a preallocated bf16 ring (ppo.RolloutMemory) instead of 200 x 31 MB fp32 clones;
  * evaluation packs several clips into one actor forward and computes all NDCG values in one launch (ppo.evaluate);
  * the best checkpoint is written asynchronously (checkpoint.save_model), same file format."""
import torch

from .. import checkpoint, data, ppo, runtime
from . import common


def main(argv=None):
    args, vit_args, num_tasks, global_rank = common.prologue("ppo", argv)
    model = ppo.ActorCritic(args, vit_args)
    reward_model = ppo.Reward(args, vit_args)
    runtime.load_strict_or_init(args.pretrained_model_path, model.actor)         # stage-1 checkpoint
    runtime.load_strict_or_init(args.reward_model_path, model.critic)            # stage-2 checkpoint (critic init)
    runtime.load_strict_or_init(args.reward_model_path, reward_model)
    if args.is_master:
        args.logger = runtime.init_logger(args)
    model = model.to(args.device)
    reward_model = reward_model.to(args.device)
    model.eval(); reward_model.eval()

    trainset = data.PpoPairs(args, args.train_path, is_train=True)
    valset = data.PpoPairs(args, args.dev_path, is_train=False)
    train_loader = data.get_dataloader(args, trainset, num_tasks, global_rank, is_train=True)
    val_loader = data.get_dataloader(args, valset, num_tasks, global_rank, is_train=False)
    instances_num, batch_size = len(trainset), args.batch_size
    args.train_steps = int(instances_num * args.epochs_num / batch_size) + 1
    if args.is_master:
        args.logger.info("Batch size: {}".format(batch_size))
        args.logger.info("The number of training instances: {}".format(instances_num))
    args.fc1_grad_bf16 = True                  # one backward per optimizer step: bf16 side buffer for out_layer.fc1
    optimizer, critic_optimizer, scheduler, critic_scheduler = ppo.build_optimizer(args, model)
    sync = common.grad_sync_for(num_tasks)
    if sync is not None:
        for m, o in ((model.actor, optimizer), (model.critic, critic_optimizer)):
            sync.broadcast_params(m)
            sync.attach(m, o)
        sync.broadcast_params(reward_model)
    args.model = model
    best_result, step, time, cycle, ragged = 0.0, 0, 0, None, []
    if args.is_master:
        args.logger.info("Start training.")
    for epoch in range(1, args.epochs_num):
        trainset = data.PpoPairs(args, args.train_path, is_train=True)           # new random pairs every epoch
        train_loader = data.get_dataloader(args, trainset, num_tasks, global_rank, is_train=True)
        train_loader.sampler.set_epoch(epoch)
        for text_emb, img_emb, tgts in train_loader:
            text, img, tgt = common.to_device(args, text_emb, img_emb, tgts)
            full = text.shape[0] == batch_size
            if cycle is None and full and args.max_timesteps == 1:
                cycle = ppo.GraphedCycle(args, model, reward_model, optimizer, critic_optimizer,
                                         args.update_timesteps * args.max_timesteps, batch_size, text.shape[1],
                                         text.shape[2], img.shape[1], text.shape[3], grad_sync=sync)
            graphed = full and args.max_timesteps == 1                       # ppo.sh:35
            state = None
            for timestep in range(args.max_timesteps):
                time += 1
                if graphed:
                    cycle.rollout(text, img, tgt)
                else:
                    # eager path: a ragged last batch of an epoch (drop_last=False), or max_timesteps > 1, where
                    # the reference feeds the previous next_state back as the critic's state (:847-850)
                    entry = ppo.rollout(model, reward_model, text, img.unsqueeze(1), tgt, state=state)
                    state = entry[1]
                    ragged.append(entry)
                if time % args.update_timesteps == 0:
                    stats = _update(args, model, optimizer, critic_optimizer, scheduler, critic_scheduler, cycle,
                                    ragged, sync)
                    step += 1
                    _report(args, stats, step)
                    with common.replicated(sync, model.actor):
                        result = ppo.evaluate(args, val_loader, step, split="val", num_tasks=num_tasks)
                    improved = args.is_master and result > best_result
                    if common.save_if_best(args, sync, ((model.actor, optimizer), (model.critic, critic_optimizer)),
                                           improved, model, args.output_model_path) and args.is_master:
                        best_result = result
                        args.logger.info("Best val indicator until now!")
    checkpoint.wait()
    if torch.distributed.is_initialized():
        torch.distributed.barrier()


def _update(args, model, optimizer, critic_optimizer, scheduler, critic_scheduler, cycle, ragged, sync):
    """train_model over everything collected since the last update (finetune/ppo.py:885-892)."""
    if not ragged and cycle is not None:
        return cycle.update(scheduler, critic_scheduler)
    memories = (list(cycle.memory) if cycle is not None else []) + ragged
    model.train()
    stats = ppo.train_model(args, model, optimizer, critic_optimizer, scheduler, critic_scheduler, memories, 0, sync)
    model.eval()
    if cycle is not None:
        cycle.memory.clear()
    ragged.clear()
    return stats


def _report(args, stats, step):
    if not args.is_master:
        return
    p_loss, v_loss, kl, old_v, v, r_ori, r, adv, rank_loss, ent = stats
    args.logger.info(f"Training step: {step}")
    # the reference pairs the names "Rank Loss" / "Advantages" with rank_loss / advantages in this order (:894-897)
    for name, val in zip(["Policy loss", "Critic Loss", "KL Penalty", "Old Values", "Values", "Rewards Ori", "Reward",
                          "Rank Loss", "Advantages", "Entropy"],
                         [p_loss, v_loss, kl, old_v, v, r_ori, r, rank_loss, adv, ent]):
        args.logger.info(f"{name}: {val}")
    args.logger.info("\nVal set evaluation.")


if __name__ == "__main__":
    main()
