"""TEST INFRASTRUCTURE ONLY — import the UNMODIFIED reference (/root/reference) on CPU.

Shims: cwd = reference root (tencentpretrain/utils/constants.py:4 opens a relative
path), sys.path += [ref, ref/finetune], an empty `h5py` module, `Tensor.cuda` = identity.
Only usable where /root/reference exists (the build container); GPU-box tests use the committed
golden vectors instead.
"""
import contextlib
import importlib
import os
import sys
import types

REF = os.environ.get("LR2_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF, "finetune"))


@contextlib.contextmanager
def _cwd(path):
    old = os.getcwd()
    os.chdir(path)
    try:
        yield
    finally:
        os.chdir(old)


_loaded = {}


def load(name):
    """Import a reference module by name, e.g. 'ppo', 'xit', 'ndcg', 'ppo_trad', 'reward_pair_dataloader'."""
    if name in _loaded:
        return _loaded[name]
    if not available():
        raise RuntimeError("reference tree not present")
    import torch
    sys.dont_write_bytecode = True
    for p in (REF, os.path.join(REF, "finetune")):
        if p not in sys.path:
            sys.path.insert(0, p)
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    with _cwd(REF):
        mod = importlib.import_module(name)
    _loaded[name] = mod
    return mod


def ppo_args(seq_length=196, max_imgs=16, visual_feat_dim=768, mode="reg"):
    import argparse
    return argparse.Namespace(mode=mode, labels_num=3, seq_length=seq_length, max_imgs=max_imgs,
                              visual_feat_dim=visual_feat_dim)
