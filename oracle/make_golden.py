"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.pt by running the UNMODIFIED reference modules
(/root/reference/{models,spectral_normalization,coordconv,spatial_recurrent}.py under oracle/shims.py) on seeded
inputs, fp32, CPU.  Run in the build container (the GPU box has no /root/reference):

    python oracle/make_golden.py

The fixtures hold inputs, outputs and *summaries* of weights/gradients (norm, sum and a fixed strided sample) so
they stay small; weights themselves are regenerated from the seed by oracle.restated.init_* in the reference's
construction order and verified against the stored summaries before any comparison.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import restated as R  # noqa: E402
from oracle import shims  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

# (name, C, H, W, A, R, B, Hn)
CONFIGS = [
    ("minipacman", 3, 15, 19, 5, 2, 2, 5),
    ("pong64", 3, 16, 16, 4, 1, 2, 4),       # 64x64 family at a size the CPU finishes quickly
    ("sc2", 4, 12, 10, 4, 2, 3, 4),
]


def summarize(t, samples=257):
    t = t.detach().float().flatten()
    idx = torch.linspace(0, t.numel() - 1, min(samples, t.numel())).long()
    return {"numel": t.numel(), "norm": t.norm().item(), "sum": t.double().sum().item(), "idx": idx,
            "val": t[idx].clone()}


def build_reference_nets(mods, C, A, Rw, seed=0):
    """Construct the reference nets in main.py:73-77 order under one seed."""
    torch.manual_seed(seed)
    with shims.cpu_cuda_noop():
        enc = mods["models"].Encoder(16, C)
        dec = mods["models"].Decoder(16, C)
        rew = mods["models"].RewardPredictor(16, Rw)
        tr = mods["models"].Transition(16, A)
    shims.apply_legacy_circular(tr)
    return {"encoder": enc, "decoder": dec, "reward_predictor": rew, "transition": tr}


def restated_nets(C, A, Rw, seed=0):
    torch.manual_seed(seed)
    return {"encoder": R.init_encoder(16, C), "decoder": R.init_decoder(16, C),
            "reward_predictor": R.init_reward_predictor(16, Rw), "transition": R.init_transition(16, A)}


def reference_step(mods, nets, states, rewards, dones, actions, A, theta, uniforms, cf, cf_indices, cf_perm, cf_h):
    """main.py:155-285 executed with the reference modules themselves (loop body transcribed 1:1)."""
    it = iter(uniforms)
    enc, dec, rew, tr = nets["encoder"], nets["decoder"], nets["reward_predictor"], nets["transition"]
    for n in nets.values():
        n.train()
        n.zero_grad()
    F = torch.nn.functional
    with shims.injected_bernoulli(lambda shape: next(it)), shims.cpu_cuda_noop():
        z = enc(states[:, 0:3])
        z_orig = z.clone()
        B = states.shape[0]
        active_mask = torch.ones(B)
        loss = 0
        terms = {}
        for t in range(1, states.shape[1] - 1):
            active_mask = active_mask * (1 - dones[:, t])
            expected_reward = rew(z)
            rd = torch.mean(torch.mean((expected_reward - rewards[:, t]) ** 2, dim=1) * active_mask)
            terms[f"Rd Loss t={t}"] = rd
            loss = loss + theta * 1e-3 * rd
            predicted = torch.sigmoid(dec(z))
            rl = F.binary_cross_entropy(predicted, states[:, t], reduction="none").mean(-1).mean(-1).mean(-1)
            rec = torch.mean(rl * active_mask)
            terms[f"Reconstruction t={t}"] = rec
            loss = loss + rec
            onehot = torch.eye(A)[actions[:, t]]
            z = tr(z, onehot)
        if cf:
            z_cf_a = z.clone()
            z_cf_b = z_orig
            unswapped = torch.ones((B, 16))
            for i in range(B):
                ia, ib = int(cf_indices[i][0]), int(cf_indices[i][1])
                unswapped[i, ia] = 0
                unswapped[i, ib] = 0
                z_cf_b[i, ia], z_cf_b[i, ib] = z_cf_b[i, ib], z_cf_b[i, ia]
            for t in range(1, cf_h):
                z_cf_b = tr(z_cf_b, torch.eye(A)[actions[:, t]])
            l = torch.abs(z_cf_a - z_cf_b).mean(-1).mean(-1) * unswapped
            l = 0.01 * torch.mean(l.mean(-1) * active_mask)
            loss = loss + l
            terms["CF Disentanglement Loss"] = l
            z_cf_a = z.clone()
            z_cf_b = z_orig
            cf_actions = actions[np.asarray(cf_perm)]
            for t in range(1, cf_h):
                z_cf_b = tr(z_cf_b, torch.eye(A)[cf_actions[:, t]])
            l = -torch.log(torch.abs(z_cf_a - z_cf_b).mean(-1).mean(-1).mean(-1) + 0.001)
            l = 0.01 * torch.mean(l * active_mask)
            loss = loss + l
            terms["CF Control Bias Loss"] = l
        loss.backward()
    return loss, terms, z


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    mods = shims.load_reference_modules()
    torch.set_num_threads(8)
    for name, C, H, W, A, Rw, B, Hn in CONFIGS:
        nets = build_reference_nets(mods, C, A, Rw, seed=0)
        out = {"config": dict(name=name, C=C, H=H, W=W, A=A, R=Rw, B=B, Hn=Hn, seed=0)}
        out["weights"] = {net: {k: summarize(v) for k, v in m.state_dict().items() if v.dtype.is_floating_point}
                          for net, m in nets.items()}
        states, rewards, dones, actions = R.synthetic_batch(B, Hn, C, H, W, A, Rw, seed=1234, p_done=0.2)
        out["inputs"] = dict(states=states, rewards=rewards, dones=dones, actions=torch.as_tensor(actions))

        # ---- per-module forward (eval-mode transition = deterministic threshold path)
        g = torch.Generator().manual_seed(99)
        zin = (torch.rand(B, 16, H, W, generator=g) < 0.5).float()
        zreal = torch.rand(B, 16, H, W, generator=g)
        onehot = torch.eye(A)[actions[:, 1]]
        u0 = torch.rand(B, 16, H, W, generator=g)
        mod = {}
        with torch.no_grad(), shims.cpu_cuda_noop():
            nets["encoder"].train()
            mod["encoder_z"] = nets["encoder"](states[:, 0:3])
            mod["decoder_logits"] = nets["decoder"](zreal)
            mod["decoder_logits_vis"] = nets["decoder"](zreal, visualize=True)[1]
            mod["reward"] = nets["reward_predictor"](zreal)
            nets["transition"].train()
            with shims.injected_bernoulli(lambda shape: u0):
                mod["transition_all"] = [t.clone() for t in nets["transition"](zin, onehot, return_all=True)]
            nets["transition"].eval()
            mod["transition_eval"] = nets["transition"](zreal, onehot)
            nets["transition"].train()
        mod.update(zin=zin, zreal=zreal, onehot=onehot, u0=u0)
        out["modules"] = mod
        # SN state after the calls above (u/v advanced once for encoder convs, twice for transition convs)
        out["sn_state"] = {net: {k: v.clone() for k, v in nets[net].state_dict().items()
                                 if k.endswith("weight_u") or k.endswith("weight_v")}
                           for net in ("encoder", "transition")}

        # ---- full training-step loss + gradients, fresh nets (CF losses on, counterfactual horizon 3)
        nets = build_reference_nets(mods, C, A, Rw, seed=0)
        n_trans = (Hn - 2) + 2 * 2
        uniforms = [torch.rand(B, 16, H, W, generator=g) for _ in range(n_trans)]
        cf_indices = torch.randint(16, (B, 2), generator=g)
        cf_perm = torch.randperm(B, generator=g)
        loss, terms, zfin = reference_step(mods, nets, states, rewards, dones, actions, A, 0.5, uniforms, True,
                                           cf_indices, cf_perm, 3)
        step = {"loss": loss.detach(), "terms": {k: v.detach() for k, v in terms.items()}, "z_final": zfin.detach(),
                "uniforms": uniforms, "cf_indices": cf_indices, "cf_perm": cf_perm, "theta": 0.5, "cf_horizon": 3}
        step["grads"] = {net: {k: summarize(p.grad) for k, p in m.named_parameters() if p.grad is not None}
                         for net, m in nets.items()}
        out["step"] = step
        path = os.path.join(GOLDEN_DIR, f"{name}.pt")
        torch.save(out, path)
        print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB) loss={loss.item():.6f}")

    # ---- interface-only layers: CoordConv2d, CSRN (reference coordconv.py, spatial_recurrent.py)
    torch.manual_seed(0)
    with shims.cpu_cuda_noop():
        cc = mods["coordconv"].CoordConv2d(6 + 2, 16, 3, padding=1)
        cs = mods["spatial_recurrent"].CSRN(8)
        g = torch.Generator().manual_seed(5)
        xc = torch.randn(2, 6, 12, 12, generator=g)
        xs = torch.randn(2, 8, 6, 7, generator=g) * 0.05
        with torch.no_grad():
            yc = cc(xc)
            ys = cs(xs)
    torch.save({"coordconv": {"state": cc.state_dict(), "x": xc, "y": yc},
                "csrn": {"state": cs.state_dict(), "x": xs, "y": ys}}, os.path.join(GOLDEN_DIR, "layers.pt"))
    print("wrote layers.pt")


if __name__ == "__main__":
    main()
