"""Data-parallel gradient exchange for the world-model training step (new functionality; the reference is
single-GPU).

Every loss of reference main.py is a batch mean with a fixed denominator, so with equal shards the global gradient is
the mean of the per-rank gradients: one sum-allreduce over the 1.17 M trainable floats per iteration, followed by
the same clip+Adam on every rank (clipping acts on the averaged gradient, as on one GPU).

All gradients live in flat fp32 buckets (`.grad` tensors are views into them, so there is no packing copy):
  bucket 0: reward predictor + decoder + transition - final as soon as BPTT reaches the first rollout step (3.4 MB);
  buckets 1..: the encoder, one bucket per layer in backward order (conv4, conv3, conv2, conv1) - the encoder's
               backward is the last thing of the iteration, so its gradients become final layer by layer.
Every bucket's allreduce is issued on a side stream from the gradient hook the moment its last gradient has been
accumulated: bucket 0 overlaps the whole encoder backward, each encoder layer overlaps the backward of the layers below
it, and only the last, smallest one (conv1: 27 C x 128 weights, ~14 KB) is exposed - with one encoder bucket the whole
1.3 MB exchange sat between the end of backward and the optimiser (round 1: 0.26 ms of a 10.3 ms iteration at 8 GPUs).
The optimiser waits for all of them.  Works eagerly and under CUDA-graph capture (fork/join through stream waits).
"""
import os

import torch
import torch.distributed as dist

# Diagnostic only (profiles/r02_notes.md, "rank skew"): skip the exchange, so that every rank runs at its own pace and
# the per-rank step times show how far the GPUs of one box are apart.  Training with it is wrong by construction.
_NO_EXCHANGE = os.environ.get("SCMGAN_DP_NOSYNC") == "1"


class BucketedGradSync:
    def __init__(self, trainer, process_group=None):
        self.pg = process_group
        self.world_size = dist.get_world_size(process_group)
        early, late = [], {}
        names = {}
        for name, net in getattr(trainer, "nets", {}).items():
            for pn, p in net.named_parameters():
                names[id(p)] = pn
        for (p, _), ni in zip(trainer.groups, trainer.net_of):
            if trainer.NET_ORDER[ni] == "encoder":
                # one bucket per encoder layer ("conv3.module.weight_bar" -> "conv3"); unknown names share one bucket
                late.setdefault(names.get(id(p), "").split(".")[0], []).append(p)
            else:
                early.append(p)
        groups = [early] + [late[k] for k in sorted(late, reverse=True)]   # conv4, conv3, ... = backward order
        self.buckets = []
        for params in groups:
            if not params:
                continue
            n = sum(p.numel() for p in params)
            flat = torch.zeros(n, dtype=torch.float32, device=params[0].device)
            off = 0
            for p in params:
                p.grad = flat[off:off + p.numel()].view_as(p)
                off += p.numel()
            self.buckets.append({"flat": flat, "params": params, "pending": 0})
        self.cuda = self.buckets[0]["flat"].is_cuda
        self.side = torch.cuda.Stream() if self.cuda else None
        self._bucket_of = {id(p): bi for bi, b in enumerate(self.buckets) for p in b["params"]}
        self._armed = False
        trainer.sync = self
        trainer.world_size = self.world_size
        if hasattr(trainer, "register_sinks"):
            trainer.register_sinks()

    def zero(self):
        for b in self.buckets:
            b["flat"].zero_()

    def arm(self, profile):
        """Call right before loss.backward(); profile: {param id: accumulations per iteration} (Trainer._profile)."""
        self._armed = True
        for b in self.buckets:
            b["pending"] = sum(profile.get(id(p), 0) for p in b["params"])

    def on_grad(self, p):
        if not self._armed:
            return
        b = self.buckets[self._bucket_of[id(p)]]
        b["pending"] -= 1
        if b["pending"] == 0:
            self._launch(b)

    def _launch(self, b):
        if _NO_EXCHANGE:
            return
        if not self.cuda:
            dist.all_reduce(b["flat"], group=self.pg)
            return
        self.side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.side):
            dist.all_reduce(b["flat"], group=self.pg)

    def finish(self):
        """Join the side stream.  A bucket whose parameters received no gradient this iteration holds zeros on every
        rank and is not exchanged."""
        self._armed = False
        if self.cuda and not _NO_EXCHANGE:
            torch.cuda.current_stream().wait_stream(self.side)
