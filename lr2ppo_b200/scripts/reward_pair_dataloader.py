"""Stage 2 -- `sh reward_pair_dataloader.sh NAME` (finetune/reward_pair_dataloader.py:437-596): pairwise reward model.
Every sample holds two tags of a clip and two 4-slot orderings (chosen / reject); one step = two forwards, hinge
relu(1 - (c - r)).mean(), one backward, AdamW, schedule.  Pair accuracy on the validation triples every
--report_steps batches, best checkpoint kept."""
import torch
import torch.distributed as dist

from .. import checkpoint, data, models, runtime, stages
from . import common


def main(argv=None):
    args, vit_args, num_tasks, global_rank = common.prologue("reward_pair_dataloader", argv)
    model = models.PairClassifier(args, vit_args)
    runtime.load_towers_or_init(args, model)
    if args.is_master:
        args.logger = runtime.init_logger(args)
    model = model.to(args.device)
    trainset = data.RewardPairs(args, args.train_path, is_train=True)
    valset = data.RewardPairs(args, args.dev_path, is_train=False)
    train_loader = data.get_dataloader(args, trainset, num_tasks, global_rank, is_train=True)
    val_loader = data.get_dataloader(args, valset, num_tasks, global_rank, is_train=False,
                                     eval_batch_size=args.batch_size)
    instances_num, batch_size = len(trainset), args.batch_size
    args.train_steps = int(instances_num * args.epochs_num / batch_size) + 1
    if args.is_master:
        args.logger.info("Batch size: {}".format(batch_size))
        args.logger.info("The number of training instances: {}".format(instances_num))
    args.fc1_grad_bf16, args.fc1_passes = True, 2      # one dY^T X GEMM over both backward passes, bf16 gradient
    optimizer, scheduler = stages.build_optimizer(args, model)
    sync = common.grad_sync_for(num_tasks)
    if sync is not None:
        sync.broadcast_params(model)
        sync.attach(model, optimizer)
    args.model = model
    total_loss, total_acc, total_cnt, best_acc, step = 0.0, 0.0, 0, 0.0, 0
    if args.is_master:
        args.logger.info("Start training.")
    for epoch in range(1, args.epochs_num + 1):
        train_loader.sampler.set_epoch(epoch)
        model.train()
        for i, (text_emb, img_emb, tgts, chosen, reject) in enumerate(train_loader):
            text, img, tgt, ch, rj = common.to_device(args, text_emb, img_emb, tgts, chosen, reject)
            loss, acc = stages.reward_train_model(args, model, optimizer, scheduler, text, img.unsqueeze(1), tgt, ch,
                                                  rj, grad_sync=sync)
            both = torch.stack([loss.detach(), acc.detach()]) / dist.get_world_size()
            dist.all_reduce(both)                      # one packed all-reduce instead of two
            total_loss += both[0].item()
            total_acc += both[1].item()
            total_cnt += 1
            step += 1
            if (i + 1) % args.report_steps == 0:
                if args.is_master:
                    args.logger.info("Epoch id: {}, Training steps: {}, Avg loss: {:.3f}, Acc: {:.3f}".format(
                        epoch, i + 1, total_loss / total_cnt, total_acc / total_cnt))
                    args.logger.info("Val set evaluation.")
                total_loss, total_acc, total_cnt = 0.0, 0.0, 0
                if args.is_master:
                    args.logger.info("Evaluating...")
                with common.replicated(sync, model):
                    acc_val = float(stages.reward_evaluate(args, model, val_loader, num_tasks=num_tasks))
                if args.is_master:
                    args.logger.info(f"val accuracy: {acc_val:.4f}")
                if common.save_if_best(args, sync, ((model, optimizer),), args.is_master and acc_val > best_acc, model,
                                       args.output_model_path) and args.is_master:
                    best_acc = acc_val
                    args.logger.info("Best Acc until now!\n")
                if args.is_master:
                    args.logger.info("Best Acc: {}".format(best_acc))
                model.train()
    checkpoint.wait()
    dist.barrier()


if __name__ == "__main__":
    main()
