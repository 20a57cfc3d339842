"""Asynchronous checkpointing and true resume.

The reference saves with a blocking `torch.save(model.state_dict(), path)` (tencentpretrain/model_saver.py:4-11,
called at finetune/ppo.py:914 whenever the validation NDCG improves): 2 GB per model are copied to the host and
serialised while the GPU idles (5-13 s stalls in the reference logs), and nothing but the weights is kept, so a run
cannot be resumed.

`save_model(model, path)` keeps the reference's name, argument order and on-disk format (a plain fp32 state_dict with
the reference's key names, loadable with `torch.load` + `load_state_dict(strict=True)`), but returns as soon as the
copies are enqueued: every CUDA tensor is first snapshotted device-to-device into a reusable staging buffer ON THE
CURRENT STREAM (about 1 ms per 6 GB at HBM speed; stream order guarantees the snapshot is taken before any later
in-place parameter / moment update, so a save can never be torn between step N and N+1), then copied from the staging
buffer to reusable pinned host buffers on a side stream while training continues, and a background thread serialises
them and renames the file into place.  `save_training_state` / `load_training_state`
add what a resume needs: optimizer moments, scheduler state, step counter and RNG states.
"""
import os
import threading

import torch


class AsyncCheckpointer:
    def __init__(self):
        self._pinned = {}           # (name, shape, dtype) -> pinned host buffer, reused across saves
        self._staging = {}          # (name, shape, dtype, device) -> device snapshot buffer, reused across saves
        self._thread = None
        self._stream = None
        self.error = None

    def wait(self):
        """Block until the previous save is on disk (also called before buffers are reused)."""
        if self._thread is not None:
            self._thread.join()
            self._thread = None
        if self.error is not None:
            err, self.error = self.error, None
            raise err

    def _stage(self, obj, prefix=""):
        """Phase 1 (current stream): device-to-device snapshot of every CUDA tensor into a reusable staging buffer.
        Returns the same nested structure with the CUDA tensors replaced by their snapshots."""
        if torch.is_tensor(obj):
            t = obj.detach()
            if not t.is_cuda:
                return t
            key = (prefix, tuple(t.shape), t.dtype, t.device)
            buf = self._staging.get(key)
            if buf is None:
                buf = torch.empty(t.shape, dtype=t.dtype, device=t.device)
                self._staging[key] = buf
            buf.copy_(t, non_blocking=True)
            return buf
        if isinstance(obj, dict):
            return {k: self._stage(v, f"{prefix}.{k}") for k, v in obj.items()}
        if isinstance(obj, (list, tuple)):
            return type(obj)(self._stage(v, f"{prefix}[{i}]") for i, v in enumerate(obj))
        return obj

    def _snapshot(self, obj, prefix=""):
        """Phase 2 (side stream): copy every tensor of a (nested) dict / list into pinned host memory,
        asynchronously for CUDA tensors."""
        if torch.is_tensor(obj):
            t = obj.detach()
            key = (prefix, tuple(t.shape), t.dtype)
            buf = self._pinned.get(key)
            if buf is None:
                buf = torch.empty(t.shape, dtype=t.dtype)
                if torch.cuda.is_available():
                    buf = buf.pin_memory()
                self._pinned[key] = buf
            buf.copy_(t, non_blocking=t.is_cuda)
            return buf
        if isinstance(obj, dict):
            return {k: self._snapshot(v, f"{prefix}.{k}") for k, v in obj.items()}
        if isinstance(obj, (list, tuple)):
            return type(obj)(self._snapshot(v, f"{prefix}[{i}]") for i, v in enumerate(obj))
        return obj

    def release_staging(self):
        """Free the device staging buffers (they are kept between saves by default: a training state is saved
        repeatedly and 180 GB of HBM leaves room for one extra copy of the ~12 GB it holds)."""
        self.wait()
        self._staging.clear()

    def save(self, obj, path):
        """obj: a state_dict or any nested dict / list of tensors and plain Python values."""
        self.save_many([(obj, path)])

    def save_many(self, items):
        """items: [(obj, path)] snapshotted together and written in order by one background thread."""
        self.wait()
        event = None
        if torch.cuda.is_available():
            if self._stream is None:
                self._stream = torch.cuda.Stream()
            # phase 1 on the current stream: anything enqueued after this call sees the tensors already copied
            staged = [self._stage(obj, f"#{i}") for i, (obj, _) in enumerate(items)]
            self._stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._stream):
                snaps = [self._snapshot(obj, f"#{i}") for i, obj in enumerate(staged)]
                event = torch.cuda.Event()
                event.record(self._stream)
        else:
            snaps = [self._snapshot(obj, f"#{i}") for i, (obj, _) in enumerate(items)]
        paths = [path for _, path in items]

        def write():
            try:
                if event is not None:
                    event.synchronize()
                for snap, path in zip(snaps, paths):
                    tmp = f"{path}.tmp.{os.getpid()}"
                    torch.save(snap, tmp)
                    os.replace(tmp, path)     # a reader never sees a half-written file
            except Exception as e:            # surfaced by the next wait()
                self.error = e

        self._thread = threading.Thread(target=write, daemon=False)
        self._thread.start()


_DEFAULT = AsyncCheckpointer()


def save_model(model, model_path, checkpointer=None):
    """Drop-in for tencentpretrain.model_saver.save_model (same format); returns before the file is written.
    Call `wait()` (or save again) before reading the file back."""
    ck = checkpointer or _DEFAULT
    target = model.module if hasattr(model, "module") else model
    ck.save(target.state_dict(), model_path)
    return ck


def wait():
    _DEFAULT.wait()


def save_training_state(path, models, optimizers, schedulers, step, checkpointer=None, extra=None):
    """Everything a bit-exact resume needs.  models / optimizers / schedulers: dicts name -> object."""
    ck = checkpointer or _DEFAULT
    state = {
        "step": int(step),
        "models": {k: (m.module if hasattr(m, "module") else m).state_dict() for k, m in models.items()},
        "optimizers": {k: o.state_dict() for k, o in optimizers.items()},
        "schedulers": {k: s.state_dict() for k, s in schedulers.items()},
        "rng": {"cpu": torch.get_rng_state(),
                "cuda": torch.cuda.get_rng_state_all() if torch.cuda.is_available() else []},
        "extra": extra,
    }
    ck.save(state, path)
    return ck


def load_training_state(path, models, optimizers, schedulers, map_location="cpu"):
    """Restores in place and returns (step, extra).  FusedAdamW rebuilds its device tables on the next step and the
    engines re-cast their bf16 shadow weights (the parameters' version counters change)."""
    state = torch.load(path, map_location=map_location, weights_only=False)
    for k, m in models.items():
        (m.module if hasattr(m, "module") else m).load_state_dict(state["models"][k], strict=True)
    for k, o in optimizers.items():
        o.load_state_dict(state["optimizers"][k])
        for attr in ("_tables", "_hyper"):
            if hasattr(o, attr):
                getattr(o, attr).clear()
    for k, s in schedulers.items():
        s.load_state_dict(state["schedulers"][k])
    torch.set_rng_state(state["rng"]["cpu"])
    if torch.cuda.is_available() and state["rng"]["cuda"]:
        torch.cuda.set_rng_state_all(state["rng"]["cuda"])
    return state["step"], state.get("extra")


# ---------------------------------------------------------------------------------------------------------------
# Sharded training state (data-parallel runs with the row-sharded out_layer.fc1 optimizer).
#
# With dist.GradSync's row sharding each rank holds the authoritative fp32 master rows and Adam moments of only its
# 1/world of out_layer.fc1 (6 GB of state per model in total); the other rows are stale locally.  A consolidated save
# would first all-gather those 6 GB and then have rank 0 write everything.  Here every rank writes its own rows
# (world writers in parallel, no collective), rank 0 adds the replicated remainder, and the loader reassembles
# complete tensors, so a run can be resumed on ANY number of ranks.  `export_model` turns a sharded directory back
# into the reference's single fp32 state_dict file (tencentpretrain/model_saver.py:4-11 format).
#
#   dir/common.pt                    rank 0: step, world, replicated model / optimizer / scheduler state, RNG
#   dir/shard-00001-of-00004.pt      rank 1: {"rows": {model: {param: (r0, r1)}}, "param" / "exp_avg" / "exp_avg_sq"}
# ---------------------------------------------------------------------------------------------------------------
def _unwrap(m):
    return m.module if hasattr(m, "module") else m


def _state_index(optimizer, param):
    """Index of `param` in optimizer.state_dict()['state'] (torch numbers the parameters group by group)."""
    i = 0
    for group in optimizer.param_groups:
        for p in group["params"]:
            if p is param:
                return i
            i += 1
    raise KeyError("parameter is not managed by this optimizer")


def _block(desc):
    """(r0, r1) -> (0, r0, r1): a row block; (dim, lo, hi) stays (K-split out_layer.fc1: a column block)."""
    return (0,) + tuple(int(x) for x in desc) if len(desc) == 2 else tuple(int(x) for x in desc)


def _shard_name(rank, world):
    return f"shard-{rank:05d}-of-{world:05d}.pt"


def save_sharded(dirpath, models, optimizers, schedulers, step, rank, world, row_shards, checkpointer=None,
                 extra=None):
    """models / optimizers / schedulers: dicts keyed alike (e.g. "actor", "critic").  row_shards:
    {model_key: {param_name: (r0, r1)}} = the rows of each row-sharded parameter that THIS rank owns
    (`dist.GradSync.row_shards(module)`; {} for a replicated model).  Every rank calls this; no collective is issued.
    Returns the checkpointer (call .wait() before reading the files back)."""
    ck = checkpointer or _DEFAULT
    os.makedirs(dirpath, exist_ok=True)
    shard = {"step": int(step), "rank": int(rank), "world": int(world), "rows": {}, "param": {}, "exp_avg": {},
             "exp_avg_sq": {}}
    common = {"step": int(step), "world": int(world), "models": {}, "optimizers": {}, "schedulers": {},
              "sharded": {}, "extra": extra,
              "rng": {"cpu": torch.get_rng_state(),
                      "cuda": torch.cuda.get_rng_state_all() if torch.cuda.is_available() else []}}
    for key, model in models.items():
        target = _unwrap(model)
        named = dict(target.named_parameters())
        mine = (row_shards or {}).get(key, {})
        # only this rank's rows of the sharded parameters are read below, so stale foreign rows are harmless here
        # (dist.GradSync otherwise refuses state_dict() on a module it has sharded)
        target._lr2_sharded_save = bool(mine)
        try:
            sd = target.state_dict()
        finally:
            target._lr2_sharded_save = False
        opt = optimizers.get(key)
        osd = None
        if opt is not None:
            opt._lr2_sharded_save = bool(mine)
            try:
                osd = opt.state_dict()
            finally:
                opt._lr2_sharded_save = False
        for name, desc in mine.items():
            dim, r0, r1 = _block(desc)
            p = named[name]
            shard["rows"].setdefault(key, {})[name] = (dim, r0, r1)
            shard["param"].setdefault(key, {})[name] = p.detach().narrow(dim, r0, r1 - r0)
            common["sharded"].setdefault(key, {})[name] = list(p.shape)
            sd.pop(name)
            if opt is not None:
                idx = _state_index(opt, p)
                st = osd["state"].get(idx)
                if st is not None:
                    shard["exp_avg"].setdefault(key, {})[name] = st["exp_avg"].narrow(dim, r0, r1 - r0)
                    shard["exp_avg_sq"].setdefault(key, {})[name] = st["exp_avg_sq"].narrow(dim, r0, r1 - r0)
                    osd["state"][idx] = {k: v for k, v in st.items() if k not in ("exp_avg", "exp_avg_sq")}
        if rank == 0:
            common["models"][key] = sd
            if osd is not None:
                common["optimizers"][key] = osd
    if rank == 0:
        common["schedulers"] = {k: s.state_dict() for k, s in schedulers.items()}
        # common.pt is written last: its presence marks rank 0's part complete
        ck.save_many([(shard, os.path.join(dirpath, _shard_name(0, world))),
                      (common, os.path.join(dirpath, "common.pt"))])
    else:
        ck.save(shard, os.path.join(dirpath, _shard_name(rank, world)))
    return ck


def _read_sharded(dirpath, map_location="cpu"):
    common_path = os.path.join(dirpath, "common.pt")
    if not os.path.exists(common_path):
        raise FileNotFoundError(f"{dirpath}: no common.pt (incomplete or not a sharded checkpoint)")
    common = torch.load(common_path, map_location=map_location, weights_only=False)
    world, step = common["world"], common["step"]
    shards = []
    for r in range(world):
        path = os.path.join(dirpath, _shard_name(r, world))
        if not os.path.exists(path):
            raise FileNotFoundError(f"{dirpath}: shard {r} of {world} is missing")
        s = torch.load(path, map_location=map_location, weights_only=False)
        if s["step"] != step or s["world"] != world or s["rank"] != r:
            raise RuntimeError(f"{path}: belongs to step {s['step']} / world {s['world']}, expected {step} / {world}")
        shards.append(s)
    return common, shards


def _assemble(shape, pieces):
    """pieces: [((r0, r1) | (dim, lo, hi), block tensor)] -> complete tensor; the blocks must tile [0, shape[dim])
    exactly along one common dimension."""
    pieces = sorted(((_block(d), t) for d, t in pieces), key=lambda x: x[0][1])
    dims = {d[0] for d, _ in pieces}
    if len(dims) != 1:
        raise RuntimeError("shards of one parameter are split along different dimensions")
    dim = dims.pop()
    at = 0
    for (_, r0, r1), t in pieces:
        if r0 != at or r1 <= r0 or t.shape[dim] != r1 - r0:
            raise RuntimeError(f"row shards do not tile the parameter: expected a piece starting at index {at} of "
                               f"dim {dim}, got [{r0}, {r1}) with extent {t.shape[dim]}")
        at = r1
    if at != shape[dim]:
        raise RuntimeError(f"row shards cover {at} of {shape[dim]} entries of dim {dim}")
    return torch.cat([t for _, t in pieces], dim=dim).reshape(shape)


def load_sharded(dirpath, models, optimizers, schedulers, map_location="cpu"):
    """Inverse of save_sharded for any current world size: every rank reads common.pt and all shard files and gets
    COMPLETE parameters and moments (re-attach dist.GradSync afterwards to shard again).  Returns (step, extra)."""
    common, shards = _read_sharded(dirpath, map_location)
    for key, model in models.items():
        target = _unwrap(model)
        sd = dict(common["models"][key])
        named = dict(target.named_parameters())
        opt = optimizers.get(key)
        osd = common["optimizers"].get(key)
        for name, shape in common["sharded"].get(key, {}).items():
            sd[name] = _assemble(shape, [(s["rows"][key][name], s["param"][key][name]) for s in shards
                                         if name in s["rows"].get(key, {})])
            if opt is not None and osd is not None:
                idx = _state_index(opt, named[name])
                if any(name in s["exp_avg"].get(key, {}) for s in shards):
                    st = dict(osd["state"].get(idx, {}))
                    for mom in ("exp_avg", "exp_avg_sq"):
                        st[mom] = _assemble(shape, [(s["rows"][key][name], s[mom][key][name]) for s in shards])
                    osd["state"][idx] = st
        target.load_state_dict(sd, strict=True)
        if opt is not None and osd is not None:
            opt.load_state_dict(osd)
            for attr in ("_tables", "_hyper"):
                if hasattr(opt, attr):
                    getattr(opt, attr).clear()
    for k, s in schedulers.items():
        s.load_state_dict(common["schedulers"][k])
    torch.set_rng_state(common["rng"]["cpu"])
    if torch.cuda.is_available() and common["rng"]["cuda"]:
        torch.cuda.set_rng_state_all(common["rng"]["cuda"])
    return common["step"], common.get("extra")


def export_model(dirpath, key, model_path):
    """Sharded directory -> the reference's single-file fp32 state_dict of model `key`
    (tencentpretrain/model_saver.py:4-11; loadable with load_state_dict(strict=True))."""
    common, shards = _read_sharded(dirpath, "cpu")
    sd = dict(common["models"][key])
    for name, shape in common["sharded"].get(key, {}).items():
        sd[name] = _assemble(shape, [(s["rows"][key][name], s["param"][key][name]) for s in shards])
    torch.save(sd, model_path)
    return sd
