#!/usr/bin/env python
"""Synthetic LRMovieNet-shaped working directory for the drop-in stage scripts (there is no network for the real
dataset): the files a reference checkout holds next to the `.sh` launchers, in the formats of SURVEY.md §5a.

    python tools/make_synthetic_lrmovienet.py --out /tmp/lr2work --clips 6 --val-clips 4

  LRMovieNet/clean_feat.h5.d/<id>/{text_emb.npy [n_tags,196,768], img_emb.npy [1,n_img,768]}   (h5 shim layout)
  LRMovieNet/first_stage_data.json, first_second_stage_data.json          clips with tags / targets
  LRMovieNet/first_second_data_pair/first_second_data_pair_10pct.json     + "index": ordered tag-index pairs
  LRMovieNet/val_data.json, test_data.json                                + "filename", "description"
  models/xlm-roberta/base_config.json, models/vit/base-16-224_config.json tower hyper-parameters (JSON merged by
                                                                          load_hyperparam)
  models/huggingface_gpt2_vocab.txt, huggingface_gpt2_merges.txt          a tiny byte-level BPE vocabulary
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lr2ppo_b200 import h5shim  # noqa: E402
from lr2ppo_b200.tokenizers import bytes_to_unicode  # noqa: E402

ROBERTA = {"emb_size": 768, "feedforward_size": 3072, "hidden_size": 768, "hidden_act": "gelu", "heads_num": 12,
           "layers_num": 12, "max_seq_length": 514, "dropout": 0.1, "data_processor": "mlm",
           "embedding": ["word", "pos", "seg"], "encoder": "transformer", "mask": "fully_visible", "target": ["mlm"]}
VIT = {"emb_size": 768, "feedforward_size": 3072, "hidden_size": 768, "hidden_act": "gelu", "heads_num": 12,
       "layers_num": 12, "dropout": 0.1, "max_seq_length": 197, "data_processor": "vit", "embedding": ["patch", "pos"],
       "remove_embedding_layernorm": True, "encoder": "transformer", "mask": "fully_visible",
       "layernorm_positioning": "pre", "target": ["cls"], "image_height": 224, "image_width": 224, "patch_size": 16}


def make_clips(rng, n, first_id, min_tags, max_tags, with_pairs, with_meta):
    clips = []
    for c in range(n):
        n_tags = int(rng.integers(min_tags, max_tags + 1))
        targets = rng.integers(0, 3, n_tags)
        targets[:3] = [0, 1, 2]                              # every relevance class present (stage-2 validation)
        rng.shuffle(targets)
        clip = {"id": f"tt{first_id + c:07d}", "tags": [{"tag": f"tag_{c}_{t}", "target": int(targets[t])}
                                                       for t in range(n_tags)]}
        if with_pairs:
            pairs = []
            for i in range(n_tags):
                for j in range(n_tags):
                    if targets[i] > targets[j]:
                        pairs.append([i, j])                  # first ranked above second
            clip["index"] = pairs[:8]
        if with_meta:
            clip["filename"] = f"shot_{first_id + c:04d}.mp4"
            clip["description"] = f"synthetic clip {first_id + c}"
        clips.append(clip)
    return clips


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--clips", type=int, default=6)
    ap.add_argument("--val-clips", type=int, default=4)
    ap.add_argument("--min-tags", type=int, default=3)
    ap.add_argument("--max-tags", type=int, default=6)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    rng = np.random.default_rng(a.seed)
    root = os.path.join(a.out, "LRMovieNet")
    os.makedirs(os.path.join(root, "first_second_data_pair"), exist_ok=True)
    train = make_clips(rng, a.clips, 0, a.min_tags, a.max_tags, True, False)
    val = make_clips(rng, a.val_clips, 10000, a.min_tags, a.max_tags, False, True)
    for clip in train + val:
        n_img = int(rng.integers(3, 24))                     # fewer and more keyframes than max_imgs = 16
        h5shim.write_group(os.path.join(root, "clean_feat.h5"), clip["id"],
                           text_emb=rng.standard_normal((len(clip["tags"]), 196, 768)).astype(np.float32),
                           img_emb=rng.standard_normal((1, n_img, 768)).astype(np.float32))
    for name, clips in (("first_stage_data.json", train), ("first_second_stage_data.json", train),
                        ("first_second_data_pair/first_second_data_pair_10pct.json", train),
                        ("val_data.json", val), ("test_data.json", val)):
        with open(os.path.join(root, name), "w") as f:
            json.dump(clips, f)
    for sub, cfg in (("xlm-roberta/base_config.json", ROBERTA), ("vit/base-16-224_config.json", VIT)):
        p = os.path.join(a.out, "models", sub)
        os.makedirs(os.path.dirname(p), exist_ok=True)
        with open(p, "w") as f:
            json.dump(cfg, f, indent=2)
    # byte-level BPE: the 256 byte symbols + a handful of merges
    syms = list(bytes_to_unicode().values())
    merges = [("t", "a"), ("ta", "g"), ("Ġ", "t"), ("c", "l"), ("cl", "i"), ("cli", "p")]
    vocab = ["<s>", "<pad>", "</s>", "<unk>"] + syms + ["".join(m) for m in merges] + ["<mask>"]
    with open(os.path.join(a.out, "models", "huggingface_gpt2_vocab.txt"), "w", encoding="utf-8") as f:
        f.write("\n".join(vocab) + "\n")
    with open(os.path.join(a.out, "models", "huggingface_gpt2_merges.txt"), "w", encoding="utf-8") as f:
        f.write("#version: 0.2\n" + "\n".join(" ".join(m) for m in merges) + "\n")
    print(f"wrote {len(train)} train / {len(val)} val clips under {a.out}")


if __name__ == "__main__":
    main()
