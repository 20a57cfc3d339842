"""Where do the element-wise gradient errors of a bf16 pipeline come from?  This test measures, on the same tensors
and the same metric as tests/test_fusion_gpu.py, the error of STOCK PyTorch bf16 mixed precision
(torch.autocast(bfloat16) over the fp32 restatement oracle/fusion_ref.py: cuBLAS bf16 GEMMs with fp32 accumulation,
fp32 LayerNorm / softmax, activations stored in bf16) against the same restatement in fp32 (TF32 off) -- the noise floor
of storing activations in bf16 -- and puts the CUDA path's error next to it.  Both are recorded in
gpurun_out/parity_errors.json; the CUDA path must stay within 2e-2 of scale or, where bf16 storage noise itself
exceeds that, within 1.5x of what stock autocast shows on that tensor."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import fusion_ref
from tests import golden_util, parity


def _ref_grads(kind, sd32, text, img, index, gw, autocast):
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in sd32.items()}
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        if kind == "actor":
            logits = fusion_ref.actor_forward(sd, text, img)
        else:
            logits = fusion_ref.critic_forward(sd, text, img, index)
    (logits.float() * gw).sum().backward()
    return logits.detach().float(), {k: v.grad.float() for k, v in sd.items()}


@pytest.mark.parametrize("kind", ["actor", "critic"])
def test_cuda_path_error_vs_stock_autocast_bf16_noise_floor(kind):
    import argparse
    from lr2ppo_b200 import models
    torch.backends.cuda.matmul.allow_tf32 = False
    c = golden_util.FUSION_CFG
    sd_cpu = golden_util.make_state_dict(kind)
    sd32 = {k: v.cuda() for k, v in sd_cpu.items()}
    text, img, tgts, index = golden_util.make_inputs(kind)
    text, img = text.cuda(), img.cuda()
    index = None if index is None else index.cuda()
    n_out = text.shape[0] * text.shape[1] if kind == "actor" else text.shape[0]
    gw = golden_util.out_grad(kind, n_out).cuda()
    l32, g32 = _ref_grads(kind, sd32, text, img, index, gw, autocast=False)
    lac, gac = _ref_grads(kind, sd32, text, img, index, gw, autocast=True)
    a = argparse.Namespace(mode="reg", labels_num=3, seq_length=c["seq_length"], max_imgs=c["max_imgs"],
                           visual_feat_dim=c["feat"])
    model = (models.Actor if kind == "actor" else models.Critic)(a, a)
    model.load_state_dict(sd_cpu, strict=True)
    model = model.cuda().eval()
    logits = model.scores(text, img) if kind == "actor" else model(text, img, None, index)
    (logits * gw).sum().backward()
    test = f"noise floor [{kind}] bs2x2"
    parity.check(test, "logits: cuda path", parity.rel_err(logits, l32), 2e-2)
    parity.check(test, "logits: stock autocast bf16", parity.rel_err(lac, l32), 1.0)
    rms_top = max((g.double().norm() / g.numel() ** 0.5).item() for g in g32.values())
    worse = 0
    for name, p in model.named_parameters():
        ref = g32[name]
        rms = (ref.double().norm() / ref.numel() ** 0.5).item()
        if rms < 1e-4 * rms_top:
            continue                                      # mathematically-zero gradients: pure noise on both sides
        ours = parity.rel_err(p.grad, ref, floor=rms)     # FULL tensors here (the reference is computed in place)
        auto = parity.rel_err(gac[name], ref, floor=rms)
        parity.check(test, name + " [elem, full tensor] stock autocast bf16 (noise floor, not a bound)", auto, 1.0)
        parity.check(test, name + " [elem, full tensor] cuda path", ours, max(2e-2, 1.5 * auto))
        worse += ours > auto
    # not a bound, a statistic worth keeping: on how many tensors the hand-written path is noisier than autocast
    parity.check(test, "fraction of tensors where the cuda path is noisier than stock autocast", worse /
                 len(list(model.named_parameters())), 1.01)
