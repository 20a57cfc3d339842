"""h5py-compatible READ shim for `LRMovieNet/clean_feat.h5` on images without h5py / HDF5.

The stage scripts only ever do `h5py.File(path, 'r')[str(id)]['text_emb'][:]` / `['img_emb'][:][0]`
(finetune/ppo.py:65-66,120-127) and, for the MSLR files, `f[qid][()]`, `.keys()`, `len(f)`.  This module serves that
subset from a directory of .npy files laid out as  <path>.d/<group>/<dataset>.npy  (memory-mapped, so a DataLoader
worker touches only the rows it slices) and also WRITES that layout (`write_group`), which the synthetic-data
generator and the on-the-fly feature extractor use.  `File(path)` opens `<path>.d` when it exists; a real HDF5 file
needs the real h5py, which the `dropin/_shims` import hook never shadows when it is installed."""
import os

import numpy as np


class Dataset:
    def __init__(self, path):
        self._path, self._arr = path, None

    def _a(self):
        if self._arr is None:
            self._arr = np.load(self._path, mmap_mode="r")
        return self._arr

    shape = property(lambda self: self._a().shape)
    dtype = property(lambda self: self._a().dtype)

    def __getitem__(self, key):
        out = self._a()[key]
        return np.array(out) if isinstance(out, np.memmap) or out.ndim else out[()]

    def __len__(self):
        return len(self._a())

    def __array__(self, dtype=None):
        return np.asarray(self._a(), dtype=dtype)


class Group:
    def __init__(self, path):
        self._path = path

    def keys(self):
        names = []
        for n in sorted(os.listdir(self._path)):
            names.append(n[:-4] if n.endswith(".npy") else n)
        return names

    def __iter__(self):
        return iter(self.keys())

    def __len__(self):
        return len(self.keys())

    def __contains__(self, key):
        base = os.path.join(self._path, str(key))
        return os.path.isdir(base) or os.path.exists(base + ".npy")

    def __getitem__(self, key):
        base = os.path.join(self._path, str(key))
        if os.path.exists(base + ".npy"):
            return Dataset(base + ".npy")
        if os.path.isdir(base):
            return Group(base)
        raise KeyError(key)


class File(Group):
    def __init__(self, name, mode="r", **_):
        if mode != "r":
            raise OSError("lr2ppo_b200.h5shim.File is read-only; use write_group() to create data")
        root = name + ".d"
        if not os.path.isdir(root):
            raise OSError(f"{root} not found: this image has no h5py / HDF5; convert {name} with "
                          "lr2ppo_b200.h5shim.write_group or generate data with tools/make_synthetic_lrmovienet.py")
        super().__init__(root)
        self.filename = name

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def write_group(h5_path, group, **datasets):
    """Store `datasets` (name -> array) under group `group` of the shim layout of `h5_path`."""
    d = os.path.join(h5_path + ".d", str(group))
    os.makedirs(d, exist_ok=True)
    for name, arr in datasets.items():
        np.save(os.path.join(d, name + ".npy"), np.asarray(arr))


def write_dataset(h5_path, name, arr):
    """Top-level dataset (the MSLR `train.h5` layout: one float64 [20, 2+F] array per query id)."""
    os.makedirs(h5_path + ".d", exist_ok=True)
    np.save(os.path.join(h5_path + ".d", str(name) + ".npy"), np.asarray(arr))
