"""Pinned, double-buffered host->device batch feeder.

The reference moves every batch with blocking `.to(device)` calls on the compute stream right before the
forward (finetune/ppo.py:826-831) and materialises `img_emb.unsqueeze(1).repeat(1, tags, 1, 1)` on the device.
Here the copy of batch i+1 runs on a dedicated copy stream while step i computes:

    feeder = DeviceFeeder(example_batch, device)
    feeder.stage(host_batch_0)
    for i in range(n):
        feeder.next_into(static_inputs)        # stream-ordered wait + D2D into the graph's static buffers
        feeder.stage(host_batch_{i+1})         # overlaps with the step below
        step()

Two device slots; a slot is re-filled only after the consumer's copy out of it has been enqueued (event
ordering on both sides, no host synchronisation).  `StatsReader` is the matching device->host side: the step's
statistics are copied asynchronously into pinned memory and read one step late, so logging never stalls the GPU.
"""
import torch


class DeviceFeeder:
    def __init__(self, example, device, slots=2):
        """example: tuple of host tensors giving the shapes / dtypes of one batch."""
        self.device = torch.device(device)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.slots = [tuple(torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in example)
                      for _ in range(slots)]
        self.ready = [torch.cuda.Event() for _ in range(slots)]        # H2D into the slot finished
        self.consumed = [torch.cuda.Event() for _ in range(slots)]     # consumer's read of the slot enqueued
        self._staged = []                                              # FIFO of filled slot ids
        self._next_fill = 0
        self._ever_consumed = [False] * slots
        self.bytes_per_batch = sum(t.numel() * t.element_size() for t in example)

    def stage(self, host_batch):
        """Start the asynchronous copy of one batch (pinned host tensors) into the next free slot."""
        if len(self._staged) >= len(self.slots):
            raise RuntimeError("DeviceFeeder: all slots are staged; consume one first")
        s = self._next_fill
        self._next_fill = (s + 1) % len(self.slots)
        with torch.cuda.stream(self.copy_stream):
            if self._ever_consumed[s]:
                self.copy_stream.wait_event(self.consumed[s])
            for dst, src in zip(self.slots[s], host_batch):
                if not src.is_pinned():
                    raise RuntimeError("DeviceFeeder.stage needs pinned host tensors (asynchronous copies)")
                dst.copy_(src, non_blocking=True)
            self.ready[s].record(self.copy_stream)
        self._staged.append(s)

    def next_into(self, dst):
        """Copy the oldest staged batch into `dst` (device tensors, e.g. a CUDA graph's static inputs) on the
        current stream, ordered after its H2D copy; the slot becomes refillable once this copy has run."""
        if not self._staged:
            raise RuntimeError("DeviceFeeder: nothing staged")
        s = self._staged.pop(0)
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(self.ready[s])
        for d, t in zip(dst, self.slots[s]):
            d.copy_(t, non_blocking=True)
        self.consumed[s].record(cur)
        self._ever_consumed[s] = True


class StatsReader:
    """Asynchronous device->host read of a small per-step result with a one-step lag."""

    def __init__(self, numel, dtype=torch.float32, depth=2):
        self.host = [torch.empty(numel, dtype=dtype).pin_memory() for _ in range(depth)]
        self.done = [torch.cuda.Event() for _ in range(depth)]
        self._i = 0
        self._pending = []

    def push(self, dev_tensor):
        """Enqueue the copy of this step's result; returns the previous step's values (host tensor) or None."""
        k = self._i % len(self.host)
        out = None
        if len(self._pending) == len(self.host):          # the buffer we are about to overwrite: drain it first
            out = self._wait(self._pending.pop(0))
        self.host[k].copy_(dev_tensor.reshape(-1), non_blocking=True)
        self.done[k].record(torch.cuda.current_stream())
        self._pending.append(k)
        self._i += 1
        if out is None and len(self._pending) > 1:
            out = self._wait(self._pending.pop(0))
        return out

    def _wait(self, k):
        self.done[k].synchronize()
        return self.host[k]

    def flush(self):
        outs = [self._wait(k) for k in self._pending]
        self._pending = []
        return outs
