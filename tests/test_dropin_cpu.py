"""Drop-in boundary on the CPU: the flag contract, the config merge and the four LRMovieNet datasets of the stage
scripts against the golden produced by the REFERENCE's own parser lines / `load_hyperparam` / `MovieNet` classes on
the same synthetic working directory (tests/golden/dropin.json, oracle/make_golden_r2.py gen_dropin); the `.sh` reader
of dropin/launch.py; the exact import block of finetune/ppo.py against the drop-in tree; the tokenizers."""
import json
import os
import random
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "dropin.json")))


@pytest.fixture(scope="module")
def workdir(tmp_path_factory):
    path = str(tmp_path_factory.mktemp("lr2work"))
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "make_synthetic_lrmovienet.py"), "--out", path,
                           "--clips", "5", "--val-clips", "4", "--seed", "3"], stdout=subprocess.DEVNULL)
    return path


def _parse(stage, workdir):
    from lr2ppo_b200 import cli
    argv = GOLD["flags"][stage].split()
    args = cli.stage_parser(stage).parse_args(argv)
    vit = cli.vit_namespace(args)
    old = os.getcwd()
    os.chdir(workdir)
    try:
        return cli.load_hyperparam(args, argv), cli.load_hyperparam(vit, argv)
    finally:
        os.chdir(old)


@pytest.mark.parametrize("stage", ["pointwise", "reward_pair_dataloader", "ppo", "ppo_eval"])
def test_flags_and_config_merge_equal_the_reference_namespaces(stage, workdir):
    args, vit = _parse(stage, workdir)
    for mine, ref in ((vars(args), GOLD["namespaces"][stage]["args"]), (vars(vit), GOLD["namespaces"][stage]["vit_args"])):
        assert set(mine) == set(ref), (sorted(set(mine) ^ set(ref)))
        for k, v in ref.items():
            assert mine[k] == v, (stage, k, mine[k], v)


DATASETS = {"ppo": "PpoPairs", "pointwise": "PointwiseClips", "reward_pair_dataloader": "RewardPairs",
            "ppo_eval": "EvalClips"}
TRAIN = {"ppo": "LRMovieNet/first_second_stage_data.json", "pointwise": "LRMovieNet/first_stage_data.json",
         "reward_pair_dataloader": "LRMovieNet/first_second_data_pair/first_second_data_pair_10pct.json",
         "ppo_eval": "LRMovieNet/val_data.json"}


@pytest.mark.parametrize("key", sorted(GOLD["datasets"]))
def test_datasets_build_the_reference_sample_lists(key, workdir):
    from lr2ppo_b200 import data
    stage, split = key.split("/")
    args, _ = _parse(stage, workdir)
    args.is_master = False
    ref = GOLD["datasets"][key]
    old = os.getcwd()
    os.chdir(workdir)
    try:
        random.seed(11); np.random.seed(11); torch.manual_seed(11)
        is_train = split == "train"
        path = TRAIN[stage] if is_train or stage == "ppo_eval" else "LRMovieNet/val_data.json"
        ds = getattr(data, DATASETS[stage])(args, path, is_train=is_train)
        assert len(ds) == ref["len"]
        assert [s[0] for s in ds.samples] == ref["ids"]
        assert [list(map(int, s[2])) for s in ds.samples] == ref["tag_index"]
        if "chosen" in ref:
            assert [s[3][0].tolist() for s in ds.samples] == ref["chosen"]
            assert [s[3][1].tolist() for s in ds.samples] == ref["reject"]
        torch.manual_seed(5)
        item = ds[min(2, len(ds) - 1)]
        assert item[0].dtype == torch.float32 and item[1].shape == (args.max_imgs, 768)
        assert abs(float(item[0].double().sum()) - ref["item2"]["text_sum"]) < 1e-6
        assert np.allclose(item[1].double().sum(dim=1).numpy(), ref["item2"]["img"], atol=1e-9)   # same shuffle + pad
        assert item[2].tolist() == ref["item2"]["tgts"]
    finally:
        os.chdir(old)


def test_launcher_reads_flag_arrays_of_a_reference_style_script(tmp_path):
    sh = tmp_path / "ppo.sh"
    sh.write_text("""TRAIN_PATH=LRMovieNet/a.json
OUTPUT_MODEL_DIR=ppo_ckpt/$1
mkdir -p ${OUTPUT_MODEL_DIR}

train_args=(
    --train_path $TRAIN_PATH
    --output_model_path ${OUTPUT_MODEL_DIR}/finetuned_model.bin
    --exp_name $1
    --max_tags 20 # 10 # 40
)
ppo_args=(
    --update_timesteps 200
)
CUDA_VISIBLE_DEVICES=0,1,2,3 torchrun --nproc_per_node=4 --master_port 29576 finetune/ppo.py \\
                                   "${train_args[@]}" \\
                                   "${ppo_args[@]}"
""")
    sys.path.insert(0, os.path.join(ROOT, "dropin"))
    import launch
    stage, flags, mkdirs = launch.parse_script(str(sh), "exp7")
    assert stage == "finetune/ppo.py" and mkdirs == ["ppo_ckpt/exp7"]
    assert flags == ["--train_path", "LRMovieNet/a.json", "--output_model_path", "ppo_ckpt/exp7/finetuned_model.bin",
                     "--exp_name", "exp7", "--max_tags", "20", "--update_timesteps", "200"]
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "dropin", "launch.py"), str(sh), "exp7", "--gpus",
                                   "8", "--dry-run", "--", "--batch_size", "2"], text=True)
    assert "--nproc-per-node=8" in out and out.strip().endswith("--batch_size 2")
    assert os.path.join("dropin", "finetune", "ppo.py") in out


def test_reference_import_block_resolves_against_the_dropin_tree():
    code = """
import sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
from tencentpretrain.embeddings import *
from tencentpretrain.encoders import *
from tencentpretrain.utils.vocab import Vocab
from tencentpretrain.utils.constants import *
from tencentpretrain.utils import *
from tencentpretrain.utils.optimizers import *
from tencentpretrain.utils.config import load_hyperparam
from tencentpretrain.utils.seed import set_seed
from tencentpretrain.utils.logging import init_logger
from tencentpretrain.utils.misc import pooling
from tencentpretrain.model_saver import save_model
from tencentpretrain.opts import finetune_opts, tokenizer_opts, adv_opts
from tencentpretrain.model_builder import build_model
import h5py
from ndcg import AverageNDCGMeter
from xit import XiT
from misc import *
import ppo, ppo_eval, pointwise, reward_pair_dataloader
for m in (ppo, ppo_eval, pointwise, reward_pair_dataloader):
    assert callable(m.main) and m.MovieNet and m.get_dataloader
assert ppo.ActorCritic and ppo.Reward and ppo.RankLoss and ppo.train_model and ppo.evaluate and ppo.build_optimizer
assert pointwise.Classifier and reward_pair_dataloader.Classifier and str2tokenizer and str2optimizer and str2scheduler
print("ok")
""" % (os.path.join(ROOT, "dropin"), os.path.join(ROOT, "dropin", "finetune"))
    assert subprocess.check_output([sys.executable, "-c", code], text=True).strip().endswith("ok")


def test_bpe_tokenizer_round_trip(workdir):
    import argparse
    from lr2ppo_b200 import tokenizers
    a = argparse.Namespace(vocab_path=os.path.join(workdir, "models", "huggingface_gpt2_vocab.txt"),
                           merges_path=os.path.join(workdir, "models", "huggingface_gpt2_merges.txt"),
                           spm_model_path=None)
    tok = tokenizers.str2tokenizer["bpe"](a)
    pieces = tok.tokenize("a tag clip, naïve")
    assert "tag" in pieces and "clip" in pieces                       # merges applied
    assert tok.decode(pieces) == "a tag clip, naïve"
    assert all(p in tok.vocab for p in pieces) and tok.convert_ids_to_tokens(tok.convert_tokens_to_ids(pieces)) == pieces
    assert tokenizers.str2tokenizer["virtual"](a).vocab == []
    with pytest.raises(ValueError):
        tokenizers.str2tokenizer["bert"](a)
