"""LRMovieNet datasets and loaders of the four multimodal stage scripts.

On disk (unchanged): `<split>.json` = list of clips `{"id", "tags": [{"tag", "target"}], ["index": [[i, j], ...]],
["filename", "description"]}` and `LRMovieNet/clean_feat.h5` with per-clip groups `text_emb [n_tags, 196, 768]`,
`img_emb [1, n_img, 768]`.  What differs per stage is only WHICH tag subsets of a clip become samples:

  stage 1  PointwiseClips   finetune/pointwise.py:77-151    every clip once, tags padded / augmented to --max_tags
  stage 2  RewardPairs      finetune/reward_pair_dataloader.py:87-206   one sample per "index" pair, chosen / reject
                                                            4-slot orderings; validation: random triples per clip
  stage 3  PpoPairs         finetune/ppo.py:58-151          --max_tags random tag pairs per clip (targets unused)
  eval     EvalClips        finetune/ppo_eval.py:60-127     every clip with all its tags (+ the clip record)

The random draws (python `random`, `numpy.random`, `torch.randperm`) are made in the reference's order, so a run seeded
like the reference (`setup_seed(seed + rank)`) builds the same sample lists.  A sample is (text_emb [T, 196, 768] fp32,
img_emb [max_imgs, 768] fp32 -- shuffled, cyclically padded --, tgts [T] int64 [, chosen [4], reject [4] | clip]).
The per-tag repeat of img_emb (finetune/ppo.py:831) is NOT done here or on the device: the fusion engine broadcasts the
[bs, 1, I, E] tensor inside its gather kernel.
"""
import json
import os
import random

import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset
from torch.utils.data.distributed import DistributedSampler

EMBED_ROOT = "LRMovieNet"          # relative to the working directory, like the reference (finetune/ppo.py:65-66)


def open_features(path):
    """h5py.File(path, 'r') when h5py is importable (the dropin launcher puts the shim on sys.path when it is not)."""
    try:
        import h5py
        return h5py.File(path, "r")
    except ImportError:
        from . import h5shim
        return h5shim.File(path, "r")


class _LRMovieNet(Dataset):
    def __init__(self, args, path, is_train=False):
        if getattr(args, "is_master", False):
            print("Loading MovieNet dataset...")
        with open(path, "r") as f:
            self.data = json.load(f)
        print(len(self.data))
        self.features_path = os.path.join(EMBED_ROOT, "clean_feat.h5")
        self._features = None                      # opened lazily: one handle per DataLoader worker process
        self.max_imgs, self.is_train, self.max_tags = args.max_imgs, is_train, args.max_tags
        self.samples = []                          # (clip id, tag records, tag indices, extras)
        self.build()
        if getattr(args, "is_master", False):
            print("Load Embedding Done!")

    def build(self):
        raise NotImplementedError

    def __len__(self):
        return len(self.samples)

    def _clip_tensors(self, clip_id, tag_index, tags):
        if self._features is None:
            self._features = open_features(self.features_path)
        grp = self._features[f"{clip_id}"]
        text = torch.from_numpy(np.asarray(grp["text_emb"][:], dtype=np.float32))[torch.as_tensor(tag_index)]
        frames = torch.from_numpy(np.asarray(grp["img_emb"][:][0], dtype=np.float32))
        n = frames.shape[0]
        frames = frames[torch.randperm(n)]                                  # shuffle the keyframes
        if n > self.max_imgs:
            img = frames[:self.max_imgs]
        else:                                                                # cyclic pad to max_imgs
            img = frames[torch.arange(self.max_imgs) % n]
        tgts = torch.tensor([int(t["target"]) for t in tags])
        return text, img, tgts

    def __getitem__(self, i):
        clip_id, tags, tag_index, extra = self.samples[i]
        return self._clip_tensors(clip_id, tag_index, tags) + tuple(extra)


class PointwiseClips(_LRMovieNet):
    def build(self):
        for clip in self.data:
            tags = clip["tags"]
            n = len(tags)
            if not self.is_train:
                self.samples.append((clip["id"], tags, list(range(n)), ()))
                continue
            if n > self.max_tags:
                tags = tags[:self.max_tags]
                index = [i % n for i in range(self.max_tags)]
            else:
                # augmentation: fill up to max_tags by cycling over the relevant (target != 0) tags, or over all of
                # them when the clip has none
                index = list(range(n))
                pool = [i for i in range(n) if int(tags[i]["target"]) != 0] or list(range(n))
                tags = list(tags)
                for i in range(n, self.max_tags):
                    j = pool[i % len(pool)]
                    tags.append(tags[j])
                    index.append(j)
            self.samples.append((clip["id"], tags, index, ()))


def _ordered_pair(tags):
    """Two random tags of `tags`; returns (chosen 4-slot index, reject 4-slot index): the first two slots name the
    pair, the last two its order, better-or-equal target first in `chosen`."""
    idx = list(range(len(tags)))
    random.shuffle(idx)
    a, b = idx[:2]
    keep, swap = [a, b, a, b], [a, b, b, a]
    return (keep, swap) if tags[a]["target"] >= tags[b]["target"] else (swap, keep)


class RewardPairs(_LRMovieNet):
    def __init__(self, args, path, is_train=False):
        super().__init__(args, path, is_train)

    def build(self):
        if not self.is_train:
            self.max_tags = 100
        for clip in self.data:
            tags = clip["tags"]
            if self.is_train:
                # ranks come from the annotated "index" pairs (first ranked above second); targets are ignored
                for pair in clip["index"]:
                    first = np.random.random() < 0.5
                    chosen, reject = ([0, 1, 0, 1], [0, 1, 1, 0]) if first else ([1, 0, 0, 1], [1, 0, 1, 0])
                    self.samples.append((clip["id"], [tags[i] for i in pair], list(pair),
                                         (torch.tensor(chosen), torch.tensor(reject))))
                continue
            by_class = {c: [i for i, t in enumerate(tags) if int(t["target"]) == c] for c in range(3)}
            if min(len(v) for v in by_class.values()) == 0:
                continue                                                       # needs one tag of every relevance
            triples = [[by_class[c][random.randint(0, len(by_class[c]) - 1)] for c in range(3)]
                       for _ in range(self.max_tags)]
            for tri in triples:
                sub = [tags[i] for i in tri]
                chosen, reject = _ordered_pair(sub)
                self.samples.append((clip["id"], sub, tri, (torch.tensor(chosen), torch.tensor(reject))))


class PpoPairs(_LRMovieNet):
    def build(self):
        for clip in self.data:
            tags = clip["tags"]
            n = len(tags)
            if not self.is_train:
                self.samples.append((clip["id"], tags, list(range(n)), ()))
                continue
            pairs = []
            for _ in range(self.max_tags):                                    # supervision comes from the reward model
                idx = list(range(n))
                random.shuffle(idx)
                pairs.append(idx[:2])
            for pair in pairs:
                self.samples.append((clip["id"], [tags[i] for i in pair], pair, ()))


class EvalClips(_LRMovieNet):
    def build(self):
        for clip in self.data:
            self.samples.append((clip["id"], clip["tags"], list(range(len(clip["tags"]))), (clip,)))


def loader_workers():
    """The reference hard-codes num_workers=32 (finetune/ppo.py:692); LR2_NUM_WORKERS overrides (0 = in-process)."""
    return int(os.environ.get("LR2_NUM_WORKERS", "32"))


def feed_bf16():
    """LR2_FEED_BF16 (default 1): the loader workers hand out text_emb / img_emb in bf16.  Every kernel of the path
    consumes the features in bf16 -- the first device operation on a fp32 batch is the round-to-nearest cast -- so doing
    that cast in the (parallel, CPU-side) workers changes no result bit and halves the pinned-memory upload
    (SURVEY.md §8(f) 1).  0 keeps the reference's fp32 batches."""
    return os.environ.get("LR2_FEED_BF16", "1") == "1"


def _collate_bf16(batch):
    from torch.utils.data import default_collate
    out = default_collate(batch)
    return type(out)(t.to(torch.bfloat16) if torch.is_tensor(t) and t.dtype == torch.float32 else t for t in out)


def get_dataloader(args, dataset, num_tasks, global_rank, is_train=False, eval_batch_size=1):
    """DistributedSampler sharding as in every stage script (finetune/ppo.py:684-699): shuffled training shards of
    --batch_size, ordered evaluation shards of `eval_batch_size` (1 for NDCG, --batch_size for stage-2 accuracy),
    drop_last=False so all ranks iterate equally.  Batches land in pinned memory for asynchronous upload, features
    already in bf16 (feed_bf16)."""
    sampler = DistributedSampler(dataset, num_replicas=num_tasks, rank=global_rank, shuffle=is_train)
    return DataLoader(dataset=dataset, batch_size=args.batch_size if is_train else eval_batch_size, sampler=sampler,
                      num_workers=loader_workers(), drop_last=False, pin_memory=torch.cuda.is_available(),
                      collate_fn=_collate_bf16 if feed_bf16() else None)
