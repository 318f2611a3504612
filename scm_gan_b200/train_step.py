"""Host-side mirror of one training iteration of the reference (main.py:143-296), built to be CUDA-graph
capturable: no host<->device synchronisation, no per-step host uploads, no Python-side randomness inside.

  loss = sum_t [ theta*reward_coef*rewardMSE_t + BCE_t ]  (+ CF disentanglement) (+ CF action control)
  loss.backward(); clip_grad_value_(enc/trans/dec, 0.1); Adam on reward/enc/dec/trans

Differences from running main.py itself (all outside the arithmetic):
  * actions arrive once as an int64 device tensor [B, Hn] (reference: CPU one-hot + H2D every step, main.py:206);
  * counterfactual indices / the action shuffle are passed in as tensors (reference: np.random, main.py:249-250,275);
  * sigmoid + BCE + means are one fused kernel (scmgan::bce_logits) and clip + Adam is one multi-tensor kernel;
  * loss terms stay on the device (reference: ts.collect() syncs per term).
"""
import collections
import os
import sys

import torch

_DROPIN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dropin")

CF_REGULARIZATION_LAMBDA = 0.01  # reference main.py:55
# SCMGAN_SEQUENTIAL_HEADS=1: evaluate reward predictor / decoder once per rollout step like main.py's loop instead of
# once per iteration on the batch of all steps (A/B measurements; the results are the same)
SEQUENTIAL_HEADS = os.environ.get("SCMGAN_SEQUENTIAL_HEADS", "0") == "1"
CLIP_VALUE = 0.1                 # reference main.py:288-290
# SCMGAN_NO_FUSED_BCE=1: decoder logits to memory + the stand-alone loss kernel (A/B measurements; same results)
FUSED_DECODER_LOSS = os.environ.get("SCMGAN_NO_FUSED_BCE", "0") != "1"


def import_dropin_models():
    """Import the drop-in `models` module (and its siblings) the way main.py would: by bare name."""
    if _DROPIN not in sys.path:
        sys.path.insert(0, _DROPIN)
    import models  # noqa: E402
    if not os.path.abspath(models.__file__).startswith(_DROPIN):
        raise RuntimeError(f"`models` resolved to {models.__file__}, not the scm_gan_b200 drop-in")
    return models


def build_nets(color_channels, num_actions, num_rewards, latent_dim=16, seed=None):
    """Construct the four trained networks in the order of reference main.py:73-77 (same RNG consumption)."""
    m = import_dropin_models()
    if seed is not None:
        torch.manual_seed(seed)
    # construct on the CPU generator (bit-identical to the reference's seeded init), then move
    enc = m.Encoder(latent_dim, color_channels)
    dec = m.Decoder(latent_dim, color_channels)
    rew = m.RewardPredictor(latent_dim, num_rewards)
    tr = m.Transition(latent_dim, num_actions)
    return {"encoder": enc, "decoder": dec, "reward_predictor": rew, "transition": tr}


def rollout_loss(nets, states, rewards, dones, actions, *, theta, reward_coef=1e-3, truncate_bptt=False,
                 enable_disentanglement=False, enable_action_control=False, cf_now=False, counterfactual_horizon=1,
                 cf_indices=None, cf_perm=None, uniforms=None, collect=None, latent_overshooting=False,
                 td_lambda=0.9):
    """Loss of one iteration (reference main.py:155-283).  All arguments are device tensors.

    states [B,Hn,C,H,W] f32, rewards [B,Hn,R] f32, dones [B,Hn] f32, actions [B,Hn] int64.
    theta: training progress train_iter/train_iters (main.py:143), a Python float or - for CUDA-graph replay, where
    one captured graph must serve every iteration - a 0-dim device tensor read by the kernels at run time.
    cf_indices [B,2] int64, cf_perm [B] int64 (when the CF losses fire); uniforms: optional list of [B,L,H,W]
    tensors consumed by successive Transition calls (parity tests), else torch's CUDA generator is used.
    collect: optional dict receiving the named loss terms the reference hands to ts.collect() (main.py:184,196,233,
    262,283) as 0-dim device tensors (no host synchronisation).
    """
    enc, dec, rew, tr = nets["encoder"], nets["decoder"], nets["reward_predictor"], nets["transition"]
    B, Hn = states.shape[0], states.shape[1]
    if not torch.is_tensor(theta):
        theta = torch.tensor(float(theta), dtype=torch.float32, device=states.device)
    A = tr.conv1.module.weight_bar.shape[1] - tr.latent_size
    eye = torch.eye(A, dtype=torch.float32, device=states.device)
    it = iter(uniforms) if uniforms is not None else None

    onehots = eye[actions.t()]  # [Hn, B, A]: one gather for the whole rollout, a contiguous [B, A] slice per step

    def step(z, a_onehot):
        if it is not None:
            tr._uniforms = next(it)
        return tr(z, a_onehot)

    z = enc(states[:, 0:3])
    z_orig = z.clone()
    # active_mask_t = prod_{s<=t} (1 - done_s)   (main.py:178)
    masks = torch.cumprod(1.0 - dones[:, 1:], dim=1)
    terms = []  # loss terms (0-dim or 1-dim device tensors); summed by ONE reduction at the end instead of an add per term
    lo_loss = torch.zeros((), dtype=torch.float32, device=states.device) if latent_overshooting else None
    lo_z = {}
    T = Hn - 2
    do_dis = bool(enable_disentanglement and cf_now)
    do_act = bool(enable_action_control and cf_now)
    n_cf = counterfactual_horizon - 1   # extra Transition calls per counterfactual branch (main.py:255-257, 276-278)
    L = z.shape[1]
    unswapped = None
    if do_dis:  # main.py:246-253
        ar = torch.arange(B, device=z.device)
        hit = torch.zeros((B, L), dtype=torch.float32, device=z.device)
        hit.scatter_(1, cf_indices, 1.0)  # both intervened factors of every sample (graph-capturable)
        unswapped = 1.0 - hit
    if T > 0 and not latent_overshooting and not SEQUENTIAL_HEADS:
        # ---- batched path ---------------------------------------------------------------------------------------
        # (1) The reward predictor and the decoder are stateless (no spectral norm, no sampling) and nothing
        #     downstream of them feeds the rollout, so the T per-step calls of main.py:181-197 are evaluated as ONE
        #     batch of T*B latents after the rollout: same arithmetic per sample, 1/T of the launches.
        # (2) Counterfactual step j of both branches (main.py:255-257, 276-278) starts from z_orig and needs nothing
        #     from the main rollout, so it is folded into the batch of main step j: one Transition call on up to 3B
        #     samples.  Each segment keeps the spectral-norm sigma of ITS call in the reference's order (all main
        #     calls, then the disentanglement branch, then the action-control branch): the power iterations only
        #     depend on the weights, so all of them run ahead in one launch (Transition.power_iterations) and u, v end
        #     in the same state.
        if do_dis:
            # main.py:253 assigns through views: net effect z[i, idx_a] <- z[i, idx_b], in place on
            # z_orig, hence also seen by the action-control branch (main.py:272)
            z_orig[ar, cf_indices[:, 0]] = z_orig[ar, cf_indices[:, 1]]
        sig = tr.power_iterations(T + n_cf * (int(do_dis) + int(do_act)))
        row_d, row_a = T, T + (n_cf if do_dis else 0)
        zd = z_orig if do_dis else None
        za = z_orig if do_act else None
        cf_onehots = onehots[:, cf_perm] if do_act else None  # actions of another trajectory of the batch (main.py:275)
        ulist = list(uniforms) if uniforms is not None else None
        zs = []
        for t in range(1, max(T, n_cf if (do_dis or do_act) else 0) + 1):
            segs, acts, rows = [], [], []
            if t <= T:
                zs.append(z)
                segs.append(z.detach() if (truncate_bptt and t > 1) else z)   # main.py:192-193
                acts.append(onehots[t])
                rows.append(t - 1)
            if do_dis and t <= n_cf:
                segs.append(zd); acts.append(onehots[t]); rows.append(row_d + t - 1)
            if do_act and t <= n_cf:
                segs.append(za); acts.append(cf_onehots[t]); rows.append(row_a + t - 1)
            if ulist is not None:
                tr._uniforms = ulist[rows[0]] if len(rows) == 1 else torch.cat([ulist[r] for r in rows], dim=0)
            if len(segs) == 1:
                outs = [tr(segs[0], acts[0], sigma=sig[rows[0]])]
            else:
                outs = list(tr(torch.cat(segs, dim=0), torch.cat(acts, dim=0),
                               sigma=torch.stack([sig[r] for r in rows])).chunk(len(segs), dim=0))
            if t <= T:
                z = outs.pop(0)
            if do_dis and t <= n_cf:
                zd = outs.pop(0)
            if do_act and t <= n_cf:
                za = outs.pop(0)
        zcat = torch.cat(zs, dim=0)                                # [T*B, L, H, W], t-major
        mask_bt = masks[:, :T]
        rd = torch.ops.scmgan.masked_mse_seq(rew(zcat), rewards[:, 1:Hn - 1], mask_bt, reward_coef, theta)
        if FUSED_DECODER_LOSS and dec.color_channels <= 16:
            # decoder + sigmoid + BCE + means: the loss head sits in the epilogue of the decoder's last convolution
            rec = dec.pixel_loss_seq(zcat, states[:, 1:Hn - 1], mask_bt)
        else:
            rec = torch.ops.scmgan.bce_logits_seq(dec(zcat), states[:, 1:Hn - 1], mask_bt)[0]
        terms += [rd[0], rec]
        if collect is not None:
            for t in range(1, Hn - 1):
                collect[f"Rd Loss t={t}"] = rd[2][t - 1]
                collect[f"Reconstruction t={t}"] = rec[t - 1]
        mask = masks[:, Hn - 3]
        if do_dis:   # main.py:258-262
            cf = torch.ops.scmgan.cf_loss(z, zd, unswapped, mask, 0, CF_REGULARIZATION_LAMBDA)[0]
            terms.append(cf)
            if collect is not None:
                collect["CF Disentanglement Loss"] = cf
        if do_act:   # main.py:279-283
            cf = torch.ops.scmgan.cf_loss(z, za, None, mask, 1, CF_REGULARIZATION_LAMBDA)[0]
            terms.append(cf)
            if collect is not None:
                collect["CF Control Bias Loss"] = cf
        loss = torch.cat([t_.reshape(-1) for t_ in terms]).sum()
        return loss, z

    # ---- sequential path: main.py's own order, call by call (latent overshooting, horizon 2, A/B measurements) ------
    for t in range(1, Hn - 1):
        mask = masks[:, t - 1]
        expected = rew(z)
        rd = torch.ops.scmgan.masked_mse(expected, rewards[:, t], mask, reward_coef, theta)
        terms.append(rd[0])
        rec = torch.ops.scmgan.bce_logits(dec(z), states[:, t], mask)[0]
        if truncate_bptt and t > 1:
            z = z.detach()
        terms.append(rec)
        if collect is not None:
            collect[f"Rd Loss t={t}"] = rd[2]
            collect[f"Reconstruction t={t}"] = rec
        z = step(z, onehots[t])

        if latent_overshooting:  # Hafner et al., reference main.py:217-230
            lo_z[t] = enc(states[:, t - 1:t + 2])
            for t_left in range(1, t):
                lo_z[t_left] = step(lo_z[t_left], onehots[t - 1])
            for t_a in range(2, t - 1):
                lo_batch = ((lo_z[t].detach() - lo_z[t_a]) ** 2).mean(-1).mean(-1).mean(-1)
                lo_loss = lo_loss + td_lambda * torch.mean(lo_batch * mask)
    mask = masks[:, Hn - 3] if Hn > 2 else torch.ones(B, device=states.device)
    if latent_overshooting:  # main.py:232-234
        terms.append(theta * lo_loss)
        if collect is not None:
            collect["LO total"] = lo_loss

    if do_dis:  # main.py:242-262
        z_cf_b = z_orig
        # main.py:253 assigns through views: net effect z[i, idx_a] <- z[i, idx_b], in place on z_orig
        z_cf_b[ar, cf_indices[:, 0]] = z_cf_b[ar, cf_indices[:, 1]]
        for t in range(1, counterfactual_horizon):
            z_cf_b = step(z_cf_b, onehots[t])
        cf = torch.ops.scmgan.cf_loss(z, z_cf_b, unswapped, mask, 0, CF_REGULARIZATION_LAMBDA)[0]
        terms.append(cf)
        if collect is not None:
            collect["CF Disentanglement Loss"] = cf

    if do_act:  # main.py:268-283
        z_cf_b = z_orig
        cf_onehots = onehots[:, cf_perm]  # actions of another trajectory of the batch (main.py:275)
        for t in range(1, counterfactual_horizon):
            z_cf_b = step(z_cf_b, cf_onehots[t])
        cf = torch.ops.scmgan.cf_loss(z, z_cf_b, None, mask, 1, CF_REGULARIZATION_LAMBDA)[0]
        terms.append(cf)
        if collect is not None:
            collect["CF Control Bias Loss"] = cf
    if not terms:  # horizon 2: no rollout step (reference main.py:162 loop is empty)
        return torch.zeros((), dtype=torch.float32, device=states.device, requires_grad=True), z
    loss = torch.cat([t.reshape(-1) for t in terms]).sum()
    return loss, z


class Trainer:
    """zero_grad -> rollout_loss -> backward -> (gradient exchange) -> fused clip+Adam, optionally replayed as one
    CUDA graph per (horizon, cf) configuration.  Mirrors reference main.py:125-129, 150-153, 285-296."""

    NET_ORDER = ("reward_predictor", "encoder", "decoder", "transition")  # opt_pred first (main.py:292-296)

    def __init__(self, nets, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, reward_coef=1e-3, loss_kwargs=None,
                 finetune_reward=False):
        from . import kernels as K
        self.K = K
        self.nets = nets
        self.lr, self.betas, self.eps = lr, betas, eps
        self.reward_coef = reward_coef
        self.finetune_reward = finetune_reward  # --finetune-reward: only opt_pred steps (main.py:292-296)
        self.loss_kwargs = dict(loss_kwargs or {})
        self.groups = []  # (param, clip)
        self.net_of = []  # index of the owning network (= of the reference's per-network Adam instance)
        for ni, name in enumerate(self.NET_ORDER):
            clip = 0.0 if name == "reward_predictor" else CLIP_VALUE
            for p in nets[name].parameters():
                if p.requires_grad:
                    self.groups.append((p, clip))
                    self.net_of.append(ni)
        dev = self.groups[0][0].device
        self.m = [torch.zeros_like(p) for p, _ in self.groups]
        self.v = [torch.zeros_like(p) for p, _ in self.groups]
        # torch.optim.Adam counts steps per parameter and skips parameters without a gradient; gradients appear
        # per network (e.g. Transition gets none at horizon 3), so one counter per network reproduces it.
        self.step_dev = torch.zeros(len(self.NET_ORDER), dtype=torch.float32, device=dev)
        # all gradients are views into one flat buffer (one memset per iteration); the backward kernels add into
        # them directly (ops.register_grad_sink), autograd's own accumulation remains for whatever they do not cover
        self.flat_grad = torch.zeros(sum(p.numel() for p, _ in self.groups), dtype=torch.float32, device=dev)
        off = 0
        for p, _ in self.groups:
            p.grad = self.flat_grad[off:off + p.numel()].view_as(p)
            off += p.numel()
            p.register_post_accumulate_grad_hook(self._on_grad)
        self.register_sinks()
        try:  # gradients are accumulated on whichever stream runs backward (side stream during graph warm-up)
            torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
        except AttributeError:
            pass
        self._counting = False
        self._counts = {}
        self._profiles = {}   # (Hn, cf_now) -> {param id: number of gradient accumulations per iteration}
        self._graphs = {}     # (batch shape, cf_now) -> (graph, static inputs, loss); insertion order = LRU order
        self._graph_pool = torch.cuda.graph_pool_handle() if torch.cuda.is_available() else None
        self.captures = 0
        self._log_ring = torch.zeros((self.LOG_RING, self.LOG_WIDTH), dtype=torch.float32, device=dev)
        self._log_slot = torch.zeros(1, dtype=torch.int64, device=dev)
        self._log_names = {}  # (Hn, cf_now) -> term names of that configuration
        self._log_keys = collections.deque(maxlen=self.LOG_RING)  # configuration of the last iterations, oldest first
        self._log_count = 0
        self.sync = None      # dp.BucketedGradSync, attached by the data-parallel launcher
        self.world_size = 1
        self.launches_per_step = {}

    def params(self):
        return [p for p, _ in self.groups]

    def register_sinks(self):
        """(Re-)point the backward kernels at the current .grad buffers (dp.BucketedGradSync re-homes them)."""
        from . import ops
        for p, _ in self.groups:
            ops.register_grad_sink(p, p.grad, self._on_grad)

    def _on_grad(self, p):
        if self._counting:
            self._counts[id(p)] = self._counts.get(id(p), 0) + 1
        elif self.sync is not None:
            self.sync.on_grad(p)

    def _zero_grads(self):
        if self.sync is not None:
            self.sync.zero()
            return
        self.flat_grad.zero_()

    def _loss(self, batch, theta, cf_now, collect=None):
        loss, _ = rollout_loss(self.nets, batch["states"], batch["rewards"], batch["dones"], batch["actions"],
                               theta=theta, reward_coef=self.reward_coef, cf_now=cf_now,
                               cf_indices=batch.get("cf_indices"), cf_perm=batch.get("cf_perm"), collect=collect,
                               # parity tests inject the Bernoulli uniforms as one more (graph-static) input
                               # [calls, B, L, H, W]; normally the Transition draws from its device Philox stream
                               uniforms=list(batch["uniforms"].unbind(0)) if "uniforms" in batch else None,
                               **self.loss_kwargs)
        return loss

    def _profile(self, batch, theta, cf_now):
        """One dry run of forward+backward (state restored afterwards) to learn which parameters receive gradients
        in this configuration and how many times autograd accumulates into each."""
        key = (batch["states"].shape[1], bool(cf_now))
        if key in self._profiles:
            return self._profiles[key]
        snap = self._snapshot_sn()
        rng = torch.cuda.get_rng_state() if torch.cuda.is_available() else None
        self._counting, self._counts = True, {}
        try:
            self._loss(batch, theta, cf_now).backward()
        finally:
            self._counting = False
        self._restore_sn(snap)
        if rng is not None:
            torch.cuda.set_rng_state(rng)
        self._profiles[key] = dict(self._counts)
        return self._profiles[key]

    def _iteration(self, batch, theta, cf_now, profile):
        """Everything from zero_grad to the optimizer step (graph-capturable)."""
        self._zero_grads()
        terms = {}
        loss = self._loss(batch, theta, cf_now, collect=terms)
        self._log_terms((batch["states"].shape[1], bool(cf_now)), terms, loss)
        if self.sync is not None:
            self.sync.arm(profile)
        loss.backward()
        if self.sync is not None:
            self.sync.finish()
        if self.finetune_reward:
            profile = {id(p): c for (p, _), ni in zip(self.groups, self.net_of)
                       if self.NET_ORDER[ni] == "reward_predictor" and id(p) in profile for c in [profile[id(p)]]}
        live = sorted({ni for (p, _), ni in zip(self.groups, self.net_of) if id(p) in profile})
        for ni in live:
            self.step_dev[ni:ni + 1] += 1.0
        # parameters without a gradient in this configuration are skipped entirely, like torch.optim.Adam does for
        # grad=None (the unused bn_conv1 affine; Transition at horizon 3)
        chunks = [(p, p.grad, m, v, clip, self.step_dev[ni:ni + 1])
                  for (p, clip), m, v, ni in zip(self.groups, self.m, self.v, self.net_of) if id(p) in profile]
        self.K.clip_adam(chunks, self.lr, self.betas[0], self.betas[1], self.eps, 0, step_dev=self.step_dev,
                         gscale=1.0 / self.world_size)
        return loss

    # -- loss-term log (reference main.py:184,196,233,262,283 ts.collect + 297 ts.print_every(10)) ---------------------
    LOG_RING = 16    # iterations kept on the device between two read-backs
    LOG_WIDTH = 128  # named terms per iteration (2 per rollout step + LO + 2 CF + total): horizons up to 62

    def _log_terms(self, key, terms, loss):
        """Append this iteration's named loss terms to a device-side ring (graph-capturable: the slot index lives on
        the device).  The reference synchronises on every ts.collect(); here nothing leaves the device until
        read_log()."""
        names = list(terms.keys()) + ["loss"]
        assert len(names) <= self.LOG_WIDTH
        self._log_names[key] = names
        row = torch.stack([v.detach().reshape(()) for v in terms.values()] + [loss.detach().reshape(())])
        row = torch.nn.functional.pad(row, (0, self.LOG_WIDTH - row.numel())).unsqueeze(0)
        self._log_ring.index_copy_(0, self._log_slot, row)
        self._log_slot.add_(1).remainder_(self.LOG_RING)

    def read_log(self, last=10):
        """ONE device->host copy: the named loss terms of the last `last` iterations (oldest first) as a list of dicts.
        Call it every 10 iterations where the reference calls ts.print_every(10) (main.py:297)."""
        last = min(last, len(self._log_keys), self.LOG_RING)
        if last == 0:
            return []
        ring = self._log_ring.cpu()
        n = self._log_count
        out = []
        for j in range(n - last, n):
            key = self._log_keys[j - n]
            names = self._log_names[key]
            row = ring[j % self.LOG_RING]
            out.append({name: row[i].item() for i, name in enumerate(names)})
        return out

    def _snapshot_sn(self):
        out = []
        for net in self.nets.values():
            for n, p in net.named_parameters():
                if n.endswith("weight_u") or n.endswith("weight_v"):
                    out.append((p, p.detach().clone()))
        return out

    @staticmethod
    def _restore_sn(snap):
        with torch.no_grad():
            for p, val in snap:
                p.copy_(val)

    def _theta_tensor(self, theta, like):
        if torch.is_tensor(theta):
            return theta
        return torch.tensor(float(theta), dtype=torch.float32, device=like.device)

    def step(self, batch, theta, cf_now=False, use_graph=False):
        """batch: dict of device tensors; theta: training progress train_iter/train_iters (float or 0-dim tensor).
        Returns the (device) loss tensor of this iteration.  With use_graph the iteration replays ONE CUDA graph per
        (batch shape, cf_now): theta is a device scalar, not part of the graph."""
        key = (batch["states"].shape[1], bool(cf_now))
        self._log_keys.append(key)
        self._log_count += 1
        try:
            if not use_graph:
                return self._iteration(batch, self._theta_tensor(theta, batch["states"]), cf_now,
                                       self._profile(batch, theta, cf_now))
            graph, static, loss = self._graph_for(batch, cf_now)
            for k, v in batch.items():
                if static[k] is not v:
                    static[k].copy_(v, non_blocking=True)
            if torch.is_tensor(theta):
                if theta is not static["theta"]:
                    static["theta"].copy_(theta, non_blocking=True)
            elif theta is not None:
                static["theta"].fill_(float(theta))
            graph.replay()
            return loss
        finally:
            # the fused optimiser updates parameters through raw pointers (no version bump): drop the decoder's
            # folded-weight cache so that a later stand-alone forward cannot see pre-update weights
            self.nets["decoder"]._fold = None

    def _graph_for(self, batch, cf_now):
        key = (tuple(batch["states"].shape), bool(cf_now))
        g = self._graphs.get(key)
        if g is None:
            if len(self._graphs) >= self.MAX_GRAPHS:  # bounded: evict the least recently used graph
                self._graphs.pop(next(iter(self._graphs)))
            g = self._capture(batch, cf_now)
        else:
            self._graphs.pop(key)
        self._graphs[key] = g  # most recently used last
        return g

    MAX_GRAPHS = 24  # 8 horizons x {regular, CF} of the default schedule (main.py:40-41,143-145) with room to spare

    def static_inputs(self, batch, theta=None, cf_now=False):
        """Capture (if needed) and return the graph's static input tensors, so callers can fill them in place
        (`theta` included: a 0-dim tensor)."""
        return self._graph_for(batch, cf_now)[1]

    def _capture(self, batch, cf_now):
        static = {k: v.clone() for k, v in batch.items()}
        static["theta"] = torch.ones((), dtype=torch.float32, device=batch["states"].device)
        theta = static["theta"]
        profile = self._profile(static, theta, cf_now)
        # warm-up on a side stream (allocator pools, lazy init); training state is restored afterwards
        snap_sn = self._snapshot_sn()
        snap_p = [(p, p.detach().clone()) for p, _ in self.groups]
        snap_m = [t.clone() for t in self.m]
        snap_v = [t.clone() for t in self.v]
        step0 = self.step_dev.clone()
        log0 = (self._log_ring.clone(), self._log_slot.clone())
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                self._iteration(static, theta, cf_now, profile)
        torch.cuda.current_stream().wait_stream(s)

        def restore():
            with torch.no_grad():
                self._restore_sn(snap_sn)
                for p, val in snap_p:
                    p.copy_(val)
                for t, val in zip(self.m, snap_m):
                    t.copy_(val)
                for t, val in zip(self.v, snap_v):
                    t.copy_(val)
                self.step_dev.copy_(step0)
                self._log_ring.copy_(log0[0])
                self._log_slot.copy_(log0[1])
        restore()
        n0 = self.K.launch_count()
        graph = torch.cuda.CUDAGraph()
        # all graphs of this trainer draw their temporaries from one private pool: they are replayed one at a time and
        # exchange nothing through pool memory (inputs are the static tensors, the only output is the loss scalar,
        # which stays referenced), so memory is bounded by the largest graph instead of growing with every
        # (horizon, cf) configuration
        with torch.cuda.graph(graph, pool=self._graph_pool):
            loss = self._iteration(static, theta, cf_now, profile)
        self.launches_per_step[(static["states"].shape[1], bool(cf_now))] = self.K.launch_count() - n0
        self.captures += 1
        return graph, static, loss
