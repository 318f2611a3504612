"""Host -> device input pipeline for the training step (reference main.py:155-158, 206: one
blocking `torch.Tensor(...).cuda()` per array and iteration).

`InputPipeline` keeps `depth` device staging sets.  `submit(host_batch)` enqueues the copies of one batch (pinned host
memory -> staging set) on a dedicated copy stream and returns immediately; `step(...)` makes the compute stream wait
for the oldest submitted batch, moves it into the trainer's (graph-static) input tensors with device-to-device copies
(~10 us for 15.7 MB) and runs the iteration.  With depth >= 2 the PCIe transfer of batch i+1 overlaps the compute of
batch i, so the end-to-end iteration time approaches the device-resident one.
"""
import collections

import torch


def pin(batch):
    """dict of numpy arrays / CPU tensors -> dict of pinned CPU tensors with the dtypes the step expects."""
    out = {}
    for k, v in batch.items():
        t = torch.as_tensor(v)
        if t.dtype in (torch.float64, torch.float16, torch.bool) and k != "actions":
            t = t.float()
        if k in ("actions", "cf_indices", "cf_perm"):
            t = t.long()
        out[k] = t.contiguous().pin_memory()
    return out


class InputPipeline:
    def __init__(self, trainer, example, depth=2, device="cuda"):
        """example: one host batch (pinned dict) fixing shapes and dtypes."""
        self.trainer = trainer
        self.copy_stream = torch.cuda.Stream(device=device)
        self.slots = [{k: torch.empty(v.shape, dtype=v.dtype, device=device) for k, v in example.items()}
                      for _ in range(depth)]
        self.events = [torch.cuda.Event() for _ in range(depth)]
        self.free = collections.deque(range(depth))
        self.ready = collections.deque()
        self.consumed = [None] * depth  # event: the compute stream has finished reading the slot

    def submit(self, host_batch):
        """Enqueue the host->device copies of one batch; returns False when every staging set is in flight."""
        if not self.free:
            return False
        j = self.free.popleft()
        if self.consumed[j] is not None:
            self.copy_stream.wait_event(self.consumed[j])
        with torch.cuda.stream(self.copy_stream):
            for k, v in host_batch.items():
                self.slots[j][k].copy_(v, non_blocking=True)
            self.events[j].record(self.copy_stream)
        self.ready.append(j)
        return True

    def step(self, theta, cf_now=False, use_graph=True):
        """Run one training iteration on the oldest submitted batch; returns the device loss tensor."""
        if not self.ready:
            raise RuntimeError("InputPipeline.step() without a submitted batch")
        j = self.ready.popleft()
        cur = torch.cuda.current_stream()
        cur.wait_event(self.events[j])
        slot = self.slots[j]
        if use_graph:
            static = self.trainer.static_inputs(slot, theta, cf_now)
            for k, v in slot.items():
                static[k].copy_(v, non_blocking=True)
            # the slot has been read: the device-to-device copies into the graph's static inputs are ordered before
            # this point, the replay only touches the static tensors
            self._release(j, cur)
            return self.trainer.step(static, theta, cf_now=cf_now, use_graph=True)
        # eager iteration: its kernels read the slot itself (states[:, t] at every rollout step, ...), so the slot
        # may only be refilled once the whole iteration has been enqueued
        loss = self.trainer.step(slot, theta, cf_now=cf_now, use_graph=False)
        self._release(j, cur)
        return loss

    def _release(self, j, stream):
        ev = torch.cuda.Event()
        ev.record(stream)
        self.consumed[j] = ev
        self.free.append(j)
