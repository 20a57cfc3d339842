"""Stage-1 (pointwise) and stage-2 (pairwise reward model) train / eval steps with the reference's signatures.

  stage 1: finetune/pointwise.py:300-313 (train_model), :316-412 (evaluate, NDCG)
  stage 2: finetune/reward_pair_dataloader.py:347-365 (train_model), :367-412 (evaluate, pair accuracy)
"""
import torch
import torch.distributed as dist

from . import losses, models, ops
from .ndcg import AverageNDCGMeter


def _all_trainable(model):
    return all(p.requires_grad for p in model.parameters())


def build_optimizer(args, model):
    """ref: finetune/pointwise.py:274-297 == finetune/reward_pair_dataloader.py:321-344 — AdamW (no decay for
    bias / gamma / beta, correct_bias=False) + the selected schedule.  Returns (optimizer, scheduler).  The fusion
    engine's bf16 weight copies are registered with the optimizer, which refreshes them in its own pass.  With
    args.fc1_grad_bf16 the out_layer.fc1 gradient lives in the bf16 side buffer FusedAdamW reads directly;
    args.fc1_passes (stage 2: 2 -- chosen and reject forwards, one backward pass each) makes the engine compute it
    with ONE dY^T X GEMM over the rows of all passes instead of accumulating a 2 GB fp32 gradient."""
    from .optim import attach_shadows, decay_groups, make_scheduler, str2optimizer
    opt_name = getattr(args, "optimizer", "adamw")
    if opt_name not in str2optimizer:
        raise ValueError(f"optimizer {opt_name!r} is outside the LR2PPO hot path (only adamw is used by the scripts)")
    optimizer = str2optimizer[opt_name](decay_groups(model.named_parameters()), lr=args.learning_rate,
                                        correct_bias=False)
    eng = getattr(model, "_engine", None)
    if eng is not None:
        attach_shadows(eng, optimizer)
        if getattr(args, "fc1_grad_bf16", False):
            eng.enable_bf16_fc1_grad(optimizer, passes=getattr(args, "fc1_passes", 1))
    return optimizer, make_scheduler(args, optimizer)


def _step(grad_sync, model, optimizer):
    """optimizer.step(); with grad_sync (dist.GradSync attached to `model`): gradient averaging over the ranks first —
    out_layer.fc1 from all-gathered wgrad operands, everything else through the flat all-reduce bucket that runs
    under the fc1 AdamW pass (same helper as stage 3)."""
    from .ppo import _sync_and_step
    _sync_and_step(grad_sync, model, optimizer)()


def _zero_grad(model):
    """model.zero_grad() (finetune/pointwise.py:302); with persistent gradient buffers (CUDA-graph replay) the buffers
    are kept and the first write of the step overwrites them instead."""
    eng = getattr(model, "_engine", None)
    if eng is not None and eng.persistent_grads:
        eng.begin_step()
    else:
        model.zero_grad()
        if eng is not None:
            eng._fc1_pending = []


class GraphedTrainStep:
    """A stage-1 / stage-2 training step captured in a CUDA graph and replayed (the ~700 launches of a stage-2 step
    cost more host time through Python than they take on the GPU).

        step = GraphedTrainStep(lambda: reward_train_model(args, model, opt, sch, *static_batch), model, opt, sch)
        for batch in loader:
            for dst, src in zip(static_batch, batch): dst.copy_(src, non_blocking=True)
            loss, acc = step.replay()

    `fn` must read its inputs from fixed device tensors and call the scheduler itself (as the reference's train_model
    does); per replay the scheduler is stepped on the host and the optimizer's hyper-parameters are uploaded
    asynchronously, exactly like ppo.GraphedStage3Step.  Dropout seeds come from the engine's device counter, so every
    replay draws fresh masks.  With a dist.GradSync inside `fn` the NCCL calls are captured too."""

    def __init__(self, fn, model, optimizer, scheduler, warmup=3):
        self.opt, self.sch = optimizer, scheduler
        model._engine.persistent_grads = True
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.out = fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.opt.frozen_hyper = True
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = fn()

    def replay(self):
        self.opt.update_hyper()          # lr of THIS step (the schedule is stepped after the update, as in train_model)
        self.graph.replay()
        self.sch.step()
        return self.out


def pointwise_train_model(args, model, optimizer, scheduler, text_emb_batch, img_emb_batch, tgts_batch,
                          grad_sync=None):
    """loss = SmoothL1(beta 0.3)(logits, tgts); backward; AdamW; scheduler (per batch).  grad_sync=None keeps the
    reference's independent replicas; a dist.GradSync averages the gradients (north_star)."""
    _zero_grad(model)
    eng = getattr(model, "_engine", None)
    if eng is not None and getattr(model, "mode", None) == "reg" and _all_trainable(model):
        # explicit forward -> fused loss kernel (value + d loss / d logits in one launch) -> explicit backward: the
        # fusion model is ONE node anyway, so torch's autograd engine adds nothing but host time (and its stream
        # bookkeeping does not survive CUDA-graph capture of two nodes in one backward, stage 2)
        models._check_inputs(text_emb_batch, img_emb_batch)
        with torch.no_grad():
            logits, ctx = eng.forward(text_emb_batch.contiguous(), img_emb_batch.contiguous(), None,
                                      train=model.training, save=True)
            loss, dl = ops.smooth_l1_loss(logits.view(-1), tgts_batch.contiguous().view(-1), 0.3)
            eng.backward(ctx, dl)
    else:
        loss, _ = model(text_emb_batch, img_emb_batch, tgts_batch)
        loss.backward()
    _step(grad_sync, model, optimizer)
    scheduler.step()
    return loss


def reward_train_model(args, model, optimizer, scheduler, text_emb_batch, img_emb_batch, tgts_batch,
                       chosen_index_batch, reject_index_batch, margin=1.0, grad_sync=None):
    """Two forwards (chosen / reject 4-slot orderings) -> hinge relu(m - (c - r)).mean() -> one backward.
    Returns (loss, acc) like the reference.  grad_sync: see pointwise_train_model (BASELINE configs[2], data-parallel
    stage 2: both backward passes accumulate the fc1 gradient from all-gathered operands, the rest is all-reduced)."""
    _zero_grad(model)
    eng = getattr(model, "_engine", None)
    if eng is not None and _all_trainable(model):
        models._check_inputs(text_emb_batch, img_emb_batch)
        text, img = text_emb_batch.contiguous(), img_emb_batch.contiguous()
        with torch.no_grad():
            chosen, ctx_c = eng.forward(text, img, chosen_index_batch.to(torch.int64).contiguous(),
                                        train=model.training, save=True)
            reject, ctx_r = eng.forward(text, img, reject_index_batch.to(torch.int64).contiguous(),
                                        train=model.training, save=True)
            loss, acc, dc, dr = ops.pair_hinge_loss(chosen, reject, margin)
            eng.backward(ctx_c, dc)
            eng.backward(ctx_r, dr)
    else:
        chosen = model(text_emb_batch, img_emb_batch, tgts_batch, chosen_index_batch)
        reject = model(text_emb_batch, img_emb_batch, tgts_batch, reject_index_batch)
        loss, acc = losses.pair_hinge_loss(chosen, reject, margin)
        loss.backward()
    _step(grad_sync, model, optimizer)
    scheduler.step()
    return loss, acc


@torch.no_grad()
def reward_evaluate(args, model, dataloader, num_tasks=None):
    """Pair accuracy over the loader; ONE all_reduce for the whole pass (the reference issues two per batch)."""
    model.eval()
    counts = torch.zeros(2, device=args.device)
    for text_emb, img_emb, tgts, chosen_index, reject_index in dataloader:
        text = text_emb.to(args.device, non_blocking=True)
        img = img_emb.unsqueeze(1).to(args.device, non_blocking=True)      # [bs, 1, I, E]: broadcast in the gather kernel
        t = tgts.to(args.device)
        c = model(text, img, t, chosen_index.to(args.device))
        r = model(text, img, t, reject_index.to(args.device))
        counts[0] += (c > r).float().sum()
        counts[1] += c.numel()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(counts)
    return counts[0] / counts[1].clamp_min(1)


@torch.no_grad()
def pointwise_evaluate(args, model, dataloader, num_tasks=None):
    """NDCG over clips of variable length (finetune/pointwise.py:316-412, 'reg' mode): scores every clip, then
    one segmented NDCG launch and one all_gather."""
    model.eval()
    scores_l, gold_l = [], []
    for text_emb, img_emb, tgts in dataloader:
        text = text_emb.to(args.device, non_blocking=True)
        img = img_emb.unsqueeze(1).to(args.device, non_blocking=True)
        scores_l.append(model(text, img, None).view(-1))
        gold_l.append(tgts.to(args.device).view(-1))
    meter = AverageNDCGMeter()
    n, nmax = len(scores_l), max(s.numel() for s in scores_l)
    scores = torch.full((n, nmax), float("-inf"), device=args.device)
    labels = torch.zeros((n, nmax), dtype=torch.int64, device=args.device)
    lens = torch.tensor([s.numel() for s in scores_l], dtype=torch.int32, device=args.device)
    for i, (s, g) in enumerate(zip(scores_l, gold_l)):
        scores[i, :s.numel()] = s
        labels[i, :g.numel()] = g
    vals = meter.batch_ndcg(scores, labels, lens=lens)
    if num_tasks and num_tasks > 1:
        gathered = [torch.zeros_like(vals) for _ in range(num_tasks)]
        dist.all_gather(gathered, vals)
        vals = torch.cat(gathered, dim=0)
    meter.add_batch(vals)
    return meter.value()
