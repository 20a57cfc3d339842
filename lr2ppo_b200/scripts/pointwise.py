"""Stage 1 -- `sh pointwise.sh NAME` (finetune/pointwise.py:433-588): pointwise relevance regression of every
(clip, tag) item, SmoothL1(beta 0.3), AdamW + linear schedule stepped per batch, NDCG evaluation every --report_steps
batches, best checkpoint kept."""
import torch
import torch.distributed as dist

from .. import checkpoint, data, models, runtime, stages
from . import common


def main(argv=None):
    args, vit_args, num_tasks, global_rank = common.prologue("pointwise", argv)
    model = models.Classifier(args, vit_args)
    runtime.load_towers_or_init(args, model)
    if args.is_master:
        args.logger = runtime.init_logger(args)
    model = model.to(args.device)
    trainset = data.PointwiseClips(args, args.train_path, is_train=True)
    valset = data.PointwiseClips(args, args.dev_path, is_train=False)
    train_loader = data.get_dataloader(args, trainset, num_tasks, global_rank, is_train=True)
    val_loader = data.get_dataloader(args, valset, num_tasks, global_rank, is_train=False)
    instances_num, batch_size = len(trainset), args.batch_size
    args.train_steps = int(instances_num * args.epochs_num / batch_size) + 1
    if args.is_master:
        args.logger.info("Batch size: {}".format(batch_size))
        args.logger.info("The number of training instances: {}".format(instances_num))
    args.fc1_grad_bf16 = args.mode == "reg"
    optimizer, scheduler = stages.build_optimizer(args, model)
    sync = common.grad_sync_for(num_tasks)
    if sync is not None:
        sync.broadcast_params(model)
        sync.attach(model, optimizer)
    args.model = model
    total_loss, best_result, step = 0.0, 0.0, 0
    if args.is_master:
        args.logger.info("Start training.")
    for epoch in range(1, args.epochs_num + 1):
        train_loader.sampler.set_epoch(epoch)
        model.train()
        for i, (text_emb, img_emb, tgts) in enumerate(train_loader):
            text, img, tgt = common.to_device(args, text_emb, img_emb, tgts)
            loss = stages.pointwise_train_model(args, model, optimizer, scheduler, text, img.unsqueeze(1), tgt,
                                                grad_sync=sync)
            loss = loss.detach().clone()
            dist.all_reduce(loss.div_(dist.get_world_size()))
            total_loss += loss.item()
            step += 1
            if (i + 1) % args.report_steps == 0:
                dist.barrier()
                if args.is_master:
                    args.logger.info("Epoch id: {}, Training steps: {}, Avg loss: {:.3f}".format(
                        epoch, i + 1, total_loss / args.report_steps))
                    args.logger.info("Val set evaluation.")
                total_loss = 0.0
                with common.replicated(sync, model):
                    ndcg = stages.pointwise_evaluate(args, model, val_loader, num_tasks=num_tasks)
                result = float(ndcg[100000000]) if args.is_master else 0.0
                if args.is_master:
                    args.logger.info("NDCG:")
                    args.logger.info("".join("\nNDCG@{}={:.4f}".format(k, ndcg[k]) for k in sorted(ndcg.keys())))
                if common.save_if_best(args, sync, ((model, optimizer),), args.is_master and result > best_result,
                                       model, args.output_model_path) and args.is_master:
                    best_result = result
                    args.logger.info("Best NDCG until now!\n")
                if args.is_master:
                    args.logger.info("Best NDCG: {}".format(best_result))
                model.train()
    checkpoint.wait()
    dist.barrier()


if __name__ == "__main__":
    main()
