"""The drop-in boundary end to end on a B200: the reference's launcher format -> dropin/launch.py ->
dropin/finetune/<stage>.py main() under torchrun on synthetic LRMovieNet-shaped data, and `evaluate` against the output
of the reference's own evaluate functions (tests/golden/evaluate.json).

  * ppo.sh-style run: 3 cycles of 4 rollout + 4 update batches, validation NDCG after each, best checkpoint written
    in the reference's format (`actor.*` + `critic.*` fp32 keys in the reference's order and shapes);
  * ppo_eval.sh-style run on that checkpoint: strict load, NDCG, case/ppo_cases.json;
  * pointwise.sh / reward_pair_dataloader.sh-style runs: a few steps, evaluation, checkpoint."""
import argparse
import json
import logging
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import restate
from tests import golden_util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

SH = """TRAIN_PATH=LRMovieNet/{train}
DEV_PATH=LRMovieNet/val_data.json
TEST_PATH=LRMovieNet/test_data.json

OUTPUT_MODEL_DIR={stage}_ckpt/$1
mkdir -p ${{OUTPUT_MODEL_DIR}}

LOG_DIR={stage}_logs/$1
mkdir -p ${{LOG_DIR}}

train_args=(
    --train_path $TRAIN_PATH
    --dev_path $DEV_PATH
    --test_path $TEST_PATH
    --epochs_num {epochs}
    --mask fully_visible
    --output_model_path ${{OUTPUT_MODEL_DIR}}/finetuned_model.bin
    --log_path ${{LOG_DIR}}/$1.txt
    --exp_name $1
    --batch_size {bs}
    --seq_length 196
    --visual_feat_dim 768
    --max_imgs 16
    --report_steps {report}
    --mode reg
    --max_tags {max_tags} # 10 # 40
{extra_train})

{ppo_block}
text_args=(
    --vocab_path models/huggingface_gpt2_vocab.txt
    --merges_path models/huggingface_gpt2_merges.txt
    --tokenizer bpe
    --config_path models/xlm-roberta/base_config.json
    --encoder transformer
)

vit_args=(
    --vit_tokenizer virtual
    --vit_config_path models/vit/base-16-224_config.json
    --vit_encoder transformer
)

CUDA_VISIBLE_DEVICES=0,1,2,3 torchrun --nproc_per_node=4 --master_port 29576 finetune/{script}.py \\
                                   "${{train_args[@]}}" \\
{ppo_use}                                   "${{text_args[@]}}" \\
                                   "${{vit_args[@]}}"
"""
PPO_BLOCK = """ppo_args=(
{pre}    --max_timesteps 1
    --eps_clip 0.2
    --kl_div_loss_weight 0.001
    --entropy_weight 0.001
    --update_timesteps 4
    --value_clip 0.5
)
"""


@pytest.fixture(scope="module")
def work(tmp_path_factory):
    path = str(tmp_path_factory.mktemp("lr2work"))
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "make_synthetic_lrmovienet.py"), "--out", path,
                           "--clips", "6", "--val-clips", "4", "--seed", "3"], stdout=subprocess.DEVNULL)
    return path


def _launch(work, name, text, port, gpus=1, exp="t1", **extra_env):
    sh = os.path.join(work, name)
    with open(sh, "w") as f:
        f.write(text)
    env = dict(os.environ, LR2_NUM_WORKERS="2", **extra_env)
    p = subprocess.run([sys.executable, os.path.join(ROOT, "dropin", "launch.py"), sh, exp, "--gpus", str(gpus),
                        "--master-port", str(port)], cwd=work, env=env, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-6000:]
    return p


def _reference_keys(kinds):
    return [(f"{k}.{n}" if len(kinds) > 1 else n, shape) for k in kinds for n, shape in golden_util.param_specs(k)]


def test_ppo_sh_then_ppo_eval_sh_through_main(work):
    text = SH.format(train="first_second_stage_data.json", stage="ppo", epochs=2, bs=2, report=100, max_tags=4,
                     extra_train="    --critic_learning_rate 1e-3\n    --learning_rate 1e-3\n", script="ppo",
                     ppo_block=PPO_BLOCK.format(pre=""), ppo_use='                                   "${ppo_args[@]}" \\\n')
    _launch(work, "ppo.sh", text, 29611)
    log = open(os.path.join(work, "ppo_logs", "t1", "t1.txt")).read()
    # 6 clips x 4 pairs / batch 2 = 12 rollout batches = 3 cycles of update_timesteps 4
    for k in (1, 2, 3):
        assert f"Training step: {k}" in log
    assert log.count("NDCG@100000000=") == 3 and log.count("Policy loss:") == 3 and "Entropy:" in log
    assert "The number of training instances: 24" in log and "Best val indicator until now!" in log
    ckpt = os.path.join(work, "ppo_ckpt", "t1", "finetuned_model.bin")
    sd = torch.load(ckpt, map_location="cpu")
    want = _reference_keys(("actor", "critic"))
    assert list(sd.keys()) == [n for n, _ in want]                        # names AND order of the reference
    for n, shape in want:
        assert tuple(sd[n].shape) == tuple(shape) and sd[n].dtype == torch.float32, n
        assert torch.isfinite(sd[n]).all(), n
    # ---- ppo_eval.sh on that checkpoint
    text = SH.format(train="first_second_stage_data.json", stage="ppo", epochs=30, bs=24, report=100, max_tags=80,
                     extra_train="", script="ppo_eval",
                     ppo_block=PPO_BLOCK.format(pre="    --pretrained_model_path ${OUTPUT_MODEL_DIR}/finetuned_model.bin\n"),
                     ppo_use='                                   "${ppo_args[@]}" \\\n')
    p = _launch(work, "ppo_eval.sh", text, 29612)
    assert "NDCG@100000000=" in p.stderr + p.stdout + open(os.path.join(work, "ppo_logs", "t1", "t1.txt")).read()
    cases = json.load(open(os.path.join(work, "case", "ppo_cases.json")))
    val = json.load(open(os.path.join(work, "LRMovieNet", "val_data.json")))
    assert len(cases) == len(val)
    for c, v in zip(cases, val):
        assert c["id"] == [v["id"]] and c["filename"] == [v["filename"]] and c["description"] == [v["description"]]
        assert [t["target"] for t in c["tags"]] == [t["target"] for t in v["tags"]]
        assert len(c["ndcg"]) == 6 and len(c["predict"]) == len(v["tags"])
        scores = [s for _, s in c["predict"]]
        assert scores == sorted(scores, reverse=True)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_ppo_sh_on_two_gpus_with_gradient_sync(work):
    """ppo.sh on 2 ranks with LR2_GRAD_SYNC=1: the K-split out_layer.fc1 (dist.Fc1Parallel), replicated evaluation and
    the collective consolidate before rank 0 saves -- the checkpoint must again be the reference's complete fp32 file."""
    text = SH.format(train="first_second_stage_data.json", stage="ppo", epochs=2, bs=2, report=100, max_tags=4,
                     extra_train="    --critic_learning_rate 1e-3\n    --learning_rate 1e-3\n", script="ppo",
                     ppo_block=PPO_BLOCK.format(pre=""), ppo_use='                                   "${ppo_args[@]}" \\\n')
    _launch(work, "ppo2.sh", text, 29621, gpus=2, exp="t2", LR2_GRAD_SYNC="1")
    log = open(os.path.join(work, "ppo_logs", "t2", "t2.txt")).read()
    # 24 training instances over 2 ranks = 12 per rank = 6 batches of 2 = 1 full cycle of update_timesteps 4 per epoch
    assert "Training step: 1" in log and "Policy loss:" in log and "NDCG@100000000=" in log
    sd = torch.load(os.path.join(work, "ppo_ckpt", "t2", "finetuned_model.bin"), map_location="cpu")
    want = _reference_keys(("actor", "critic"))
    assert list(sd.keys()) == [n for n, _ in want]
    for n, shape in want:
        assert tuple(sd[n].shape) == tuple(shape) and sd[n].dtype == torch.float32 and torch.isfinite(sd[n]).all(), n
    # every column block of the K-split weight was updated by its owner and gathered back: no block is left at its
    # initial value on rank 0's copy (lr 1e-3 moves every element that has a gradient)
    w = sd["actor.out_layer.fc1.weight"]
    half = w.shape[1] // 2
    assert w[:, :half].abs().sum() > 0 and w[:, half:].abs().sum() > 0


@pytest.mark.parametrize("stage", ["pointwise", "reward_pair_dataloader"])
def test_stage1_and_stage2_sh_through_main(work, stage):
    train = "first_stage_data.json" if stage == "pointwise" else "first_second_data_pair/first_second_data_pair_10pct.json"
    text = SH.format(train=train, stage=stage, epochs=1, bs=2, report=2, max_tags=6, extra_train="", script=stage,
                     ppo_block="", ppo_use="")
    _launch(work, stage + ".sh", text, 29613 if stage == "pointwise" else 29614)
    log = open(os.path.join(work, f"{stage}_logs", "t1", "t1.txt")).read()
    assert "Training steps: 2" in log and "Start training." in log
    assert ("NDCG@100000000=" in log) if stage == "pointwise" else ("val accuracy:" in log)
    sd = torch.load(os.path.join(work, f"{stage}_ckpt", "t1", "finetuned_model.bin"), map_location="cpu")
    want = _reference_keys(("actor",) if stage == "pointwise" else ("reward",))
    assert list(sd.keys()) == [n for n, _ in want]
    assert all(tuple(sd[n].shape) == tuple(s) and sd[n].dtype == torch.float32 for n, s in want)


def test_evaluate_vs_the_reference_evaluate_functions(work, tmp_path):
    """finetune/ppo.py:620-681 and finetune/ppo_eval.py:401-471 were run from the imported reference on the same
    synthetic validation split and actor weights (tests/golden/evaluate.json).  Scores must agree within 2e-2 of
    scale; where the predicted order is the reference's (it is, unless two scores are closer than that tolerance) the
    per-clip NDCG lists must be BIT-identical, and in any case bit-identical to the oracle on our own order."""
    import random
    from lr2ppo_b200 import cli, data, ppo
    gold = json.load(open(os.path.join(GOLD, "evaluate.json")))
    flags = json.load(open(os.path.join(GOLD, "dropin.json")))["flags"]["ppo_eval"].split()
    args = cli.stage_parser("ppo_eval").parse_args(flags)
    old = os.getcwd()
    # the golden's working directory was generated with the same tool and seed
    wd = str(tmp_path / "w")
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "make_synthetic_lrmovienet.py"), "--out", wd,
                           "--clips", "5", "--val-clips", "4", "--seed", "3"], stdout=subprocess.DEVNULL)
    os.chdir(wd)
    os.environ["LR2_NUM_WORKERS"] = "0"
    try:
        args = cli.load_hyperparam(args, flags)
        args.labels_num, args.is_master, args.device = 3, True, torch.device("cuda")
        args.logger = logging.getLogger("test")
        model = ppo.ActorCritic(args, args)
        model.actor.load_state_dict(golden_util.make_state_dict("actor"), strict=True)
        model.critic.load_state_dict(golden_util.make_state_dict("critic"), strict=True)
        args.model = model.cuda()
        random.seed(11); np.random.seed(11); torch.manual_seed(11)
        loader = data.get_dataloader(args, data.EvalClips(args, "LRMovieNet/val_data.json"), 1, 0, is_train=False)
        torch.manual_seed(123)
        result = ppo.evaluate(args, loader, 0, split="val", num_tasks=1, cases_path="case/ppo_cases.json")
        cases = json.load(open("case/ppo_cases.json"))
    finally:
        os.chdir(old)
        os.environ.pop("LR2_NUM_WORKERS", None)
    ref_cases = gold["ppo_eval"]["cases"]
    assert len(cases) == len(ref_cases)
    same_order = True
    ks = [1, 3, 5, 10, 20, 100000000]
    for c, r in zip(cases, ref_cases):
        assert c["id"] == r["id"] and c["filename"] == r["filename"] and c["description"] == r["description"]
        assert c["tags"] == r["tags"]
        ours = {p[0]["tag"][0]: p[1] for p in c["predict"]}
        ref = {p[0]["tag"][0]: p[1] for p in r["predict"]}
        scale = max(abs(v) for v in ref.values())
        for tag, s in ref.items():
            assert abs(ours[tag] - s) < 2e-2 * scale, (tag, ours[tag], s)
        order_o = [p[0]["tag"][0] for p in c["predict"]]
        order_r = [p[0]["tag"][0] for p in r["predict"]]
        if order_o == order_r:
            assert c["ndcg"] == r["ndcg"]                                   # bit-identical fp32 values
        else:
            same_order = False
            for a, b in zip(order_o, order_r):                              # only near-ties may swap
                assert a == b or abs(ref[a] - ref[b]) < 2e-2 * scale
        labels = np.array([[t["target"] for t in c["tags"]]], dtype=np.int64)
        by_tag = {t["tag"][0]: i for i, t in enumerate(c["tags"])}
        sc = np.zeros((1, labels.shape[1]), dtype=np.float32)
        for rank, tag in enumerate(order_o):
            sc[0, by_tag[tag]] = -rank                                       # any scores with our order
        assert np.array(c["ndcg"], dtype=np.float32).tobytes() == restate.ndcg_at_k(sc, labels, ks)[0].tobytes()
    if same_order:
        assert abs(float(result) - gold["ppo_eval"]["result"]) < 1e-6 and gold["ppo"]["result"] == gold["ppo_eval"]["result"]
