"""Asynchronous checkpointing and true resume.

The reference saves with a blocking `torch.save(model.state_dict(), path)` (tencentpretrain/model_saver.py:4-11,
called at finetune/ppo.py:914 whenever the validation NDCG improves): 2 GB per model are copied to the host and
serialised while the GPU idles (5-13 s stalls in the reference logs), and nothing but the weights is kept, so a run
cannot be resumed.

`save_model(model, path)` keeps the reference's name, argument order and on-disk format (a plain fp32 state_dict with
the reference's key names, loadable with `torch.load` + `load_state_dict(strict=True)`), but returns as soon as the
device->host copies are enqueued: the tensors are snapshotted into reusable pinned buffers on a side stream and a
background thread serialises them and renames the file into place.  `save_training_state` / `load_training_state`
add what a resume needs: optimizer moments, scheduler state, step counter and RNG states.
"""
import os
import threading

import torch


class AsyncCheckpointer:
    def __init__(self):
        self._pinned = {}           # (name, shape, dtype) -> pinned host buffer, reused across saves
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

    def _snapshot(self, obj, prefix=""):
        """Copy every tensor of a (nested) dict / list into pinned host memory, asynchronously for CUDA tensors."""
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

    def save(self, obj, path):
        """obj: a state_dict or any nested dict / list of tensors and plain Python values."""
        self.wait()
        event = None
        if torch.cuda.is_available():
            if self._stream is None:
                self._stream = torch.cuda.Stream()
            self._stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._stream):
                snap = self._snapshot(obj)
                event = torch.cuda.Event()
                event.record(self._stream)
        else:
            snap = self._snapshot(obj)

        def write():
            try:
                if event is not None:
                    event.synchronize()
                tmp = f"{path}.tmp.{os.getpid()}"
                torch.save(snap, tmp)
                os.replace(tmp, path)         # a reader never sees a half-written file
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
