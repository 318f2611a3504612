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


class DeviceReplayBuffer:
    """The reference's replay buffer (envs/minipacman.py:16-27, 106-164: a list of up to REPLAY_BUFFER_LEN episodes, a
    training / testing hold-out split, `get_trajectories` sampling clips of random episodes) resident in HBM.

    `add_episode` is the only host->device traffic: one episode (<= max_len steps) is copied into a slot when the
    simulator finishes it - appended while the buffer grows, then replacing a random slot, like add_to_replay_buffer
    (envs/minipacman.py:106-113).  `get_trajectories` launches scmgan_replay_sample: every batch row is assembled from
    clips on the device (device-side Philox stream), so a training iteration needs no per-step upload at all and the
    sampling can be captured into the step's CUDA graph.  Returns device tensors
    (states [B,T,C,H,W] f32, rewards [B,T,R] f32, dones [B,T] f32 0/1, actions [B,T] int64): the tensors main.py:155-158
    builds from the numpy arrays, without the numpy detour.
    """

    def __init__(self, frame_shape, num_rewards, capacity=50, max_len=150, min_len=4, holdout=0.20, seed=0,
                 device="cuda"):
        from . import kernels as K
        self.K = K
        self.frame_shape = tuple(frame_shape)
        self.R, self.capacity, self.max_len, self.min_len = num_rewards, capacity, max_len, min_len
        self.holdout = holdout
        self.device = device
        self.host_rng = __import__("numpy").random.RandomState(seed)   # slot replacement / hold-out coin (host decisions)
        self.sets = {}
        for name in ("training", "testing"):
            self.sets[name] = {
                "frames": torch.zeros((capacity, max_len) + self.frame_shape, dtype=torch.float32, device=device),
                "rewards": torch.zeros((capacity, max_len, num_rewards), dtype=torch.float32, device=device),
                "actions": torch.zeros((capacity, max_len), dtype=torch.int32, device=device),
                "ep_len": torch.zeros(capacity, dtype=torch.int32, device=device),
                "n_filled": torch.zeros((), dtype=torch.int32, device=device),
                "count": 0,
            }
        self.rng_state = torch.tensor([seed & 0x7FFFFFFFFFFFFFFF, 0], dtype=torch.int64, device=device)
        self.copy_stream = torch.cuda.Stream(device=device) if torch.cuda.is_available() else None
        self._staged = []   # pinned host tensors kept alive until their copies have run

    def __len__(self):
        return self.sets["training"]["count"]

    def add_episode(self, states, rewards, actions, training=None):
        """states [n, C, H, W], rewards [n, R], actions [n] (numpy / CPU tensors), n >= min_len; longer episodes are cut
        at max_len (the reference's simulators stop at MAX_TRAJECTORY_LEN = 150)."""
        n = min(len(states), self.max_len)
        if n < self.min_len:
            raise ValueError(f"episode of {n} steps: get_trajectories needs at least {self.min_len}")
        if training is None:
            training = self.host_rng.random_sample() > self.holdout
        S = self.sets["training" if training else "testing"]
        slot = S["count"] if S["count"] < self.capacity else int(self.host_rng.randint(0, self.capacity))
        host = [torch.as_tensor(states[:n]).float().contiguous().pin_memory(),
                torch.as_tensor(rewards[:n]).float().reshape(n, self.R).contiguous().pin_memory(),
                torch.as_tensor(actions[:n]).to(torch.int32).contiguous().pin_memory()]
        cur = torch.cuda.current_stream()
        self.copy_stream.wait_stream(cur)   # samplers already enqueued read the old slot contents first
        with torch.cuda.stream(self.copy_stream):
            S["frames"][slot, :n].copy_(host[0], non_blocking=True)
            S["rewards"][slot, :n].copy_(host[1], non_blocking=True)
            S["actions"][slot, :n].copy_(host[2], non_blocking=True)
            S["ep_len"][slot:slot + 1].fill_(n)
            if S["count"] < self.capacity:
                S["count"] += 1
                S["n_filled"].fill_(S["count"])
        cur.wait_stream(self.copy_stream)
        self._staged = self._staged[-8:] + [host]
        return slot

    def get_trajectories(self, batch_size=8, timesteps=10, random_start=True, training=True, out=None, plan=None):
        """Device-side sampler (reference signature, envs/minipacman.py:122).  `out` (optional dict with states / rewards
        / dones / actions) receives the batch in place - e.g. the static inputs of a captured training graph."""
        S = self.sets["training" if training else "testing"]
        if S["count"] == 0:
            raise RuntimeError("replay buffer is empty")   # the reference waits for its simulator thread here
        dev = self.device
        if out is None:
            out = {"states": torch.empty((batch_size, timesteps) + self.frame_shape, dtype=torch.float32, device=dev),
                   "rewards": torch.empty((batch_size, timesteps, self.R), dtype=torch.float32, device=dev),
                   "dones": torch.empty((batch_size, timesteps), dtype=torch.float32, device=dev),
                   "actions": torch.empty((batch_size, timesteps), dtype=torch.int64, device=dev)}
        self.K.replay_sample(S["frames"], S["rewards"], S["actions"], S["ep_len"], S["n_filled"], self.rng_state,
                             out["states"], out["rewards"], out["dones"], out["actions"], random_start, plan)
        return out["states"], out["rewards"], out["dones"], out["actions"]
