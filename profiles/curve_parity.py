"""Training-curve parity (BASELINE.json north_star: "rollout-MSE curve matching over 1k steps").

Trains the B200 implementation (scm_gan_b200.Trainer) and the fp32 oracle restatement (oracle/restated.py + torch
Adam, executed with torch on the same GPU) from IDENTICAL initial weights on the SAME stream of synthetic trajectories
(scm_gan_b200.synthetic.MovingDots) for N iterations, both counterfactual losses every 5th iteration, then evaluates
both sets of learned weights with the oracle's `measure_prediction_mse` (reference main.py:784-836) on held-out
trajectories.  The Bernoulli latents use independent random streams, so the comparison is statistical: smoothed loss
curves and the rollout-MSE curve.

    python profiles/curve_parity.py --steps 1000 --out profiles/r01_curve_parity.json
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def smooth(x, w):
    x = np.asarray(x, dtype=np.float64)
    if len(x) < w:
        return x
    c = np.cumsum(np.insert(x, 0, 0.0))
    return (c[w:] - c[:-w]) / w


def run(steps, batch, horizon, seed=0, cf_h=2, use_graph=True, log=print, rng_seed=None, schedule=False):
    """schedule: the reference's horizon schedule (main.py:143-147: horizon = 3 + int((hmax - 3) * theta), hmax =
    `horizon`) instead of a fixed horizon - one CUDA graph per (horizon, cf) pair, theta exact (a device scalar)."""
    from oracle import restated as R
    from scm_gan_b200.synthetic import MovingDots
    from scm_gan_b200.train_step import Trainer, build_nets
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dev = "cuda"
    C, H, W, A, Rw = 3, 15, 19, 5, 2
    nets = build_nets(C, A, Rw, seed=seed)
    for n in nets.values():
        n.train()
    onets = {k: {n: v.detach().clone() for n, v in m.state_dict().items()} for k, m in nets.items()}
    oparams_clip, oparams_free = [], []
    for name, sd in onets.items():
        for k, v in sd.items():
            if v.dtype.is_floating_point and not (k.endswith("_u") or k.endswith("_v") or "bn_conv1" in k):
                v.requires_grad_(True)
                (oparams_free if name == "reward_predictor" else oparams_clip).append(v)
    opt = torch.optim.Adam(oparams_clip + oparams_free, lr=1e-4)
    kw = dict(enable_disentanglement=True, enable_action_control=True, counterfactual_horizon=cf_h)
    trainer = Trainer(nets, loss_kwargs=kw)
    if rng_seed is not None:  # vary only the Bernoulli streams (ours: Philox seed; oracle: torch generator)
        nets["transition"]._rng_state[0] = int(rng_seed)
        torch.manual_seed(int(rng_seed))
    src_a, src_b = MovingDots(C, H, W, A, Rw, seed=11), MovingDots(C, H, W, A, Rw, seed=11)
    g = torch.Generator().manual_seed(5)
    ours, oracle = [], []
    t_ours = t_oracle = 0.0
    for it in range(1, steps + 1):
        theta = it / steps
        hn = (3 + int((horizon - 3) * theta)) if schedule else horizon
        cf_now = (it % 5 == 0)
        cf_idx = torch.randint(16, (batch, 2), generator=g)
        cf_perm = torch.randperm(batch, generator=g)
        # ---- ours
        st, rw, dn, ac = src_a.get_trajectories(batch, hn)
        b = {"states": torch.from_numpy(st).to(dev), "rewards": torch.from_numpy(rw).to(dev),
             "dones": torch.from_numpy(dn.astype(np.float32)).to(dev), "actions": torch.from_numpy(ac).to(dev),
             "cf_indices": cf_idx.to(dev), "cf_perm": cf_perm.to(dev)}
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        # theta = train_iter / train_iters exactly (main.py:143): a device scalar, not part of the graph key
        th_q = theta
        trainer.step(b, th_q, cf_now=cf_now, use_graph=use_graph)
        if it % 10 == 0 or it == steps:   # ONE device->host copy per 10 iterations (main.py:297 ts.print_every(10))
            n_new = it - len(ours)
            ours.extend(row["loss"] for row in trainer.read_log(n_new))
        t_ours += time.perf_counter() - t0
        # ---- oracle (same data stream, same theta, same CF draws)
        st2, rw2, dn2, ac2 = src_b.get_trajectories(batch, hn)
        assert np.array_equal(st, st2)
        t0 = time.perf_counter()
        opt.zero_grad()
        ol, _, _ = R.train_step_loss(onets, b["states"], b["rewards"], b["dones"], ac2, num_actions=A, theta=th_q,
                                     cf_now=cf_now, cf_indices=cf_idx, cf_perm=cf_perm, **kw)
        ol.backward()
        torch.nn.utils.clip_grad_value_([p for p in oparams_clip if p.grad is not None], 0.1)
        opt.step()
        oracle.append(ol.item())
        t_oracle += time.perf_counter() - t0
        if it % 100 == 0:
            log(f"iter {it}: ours {np.mean(ours[-50:]):.4f}  oracle {np.mean(oracle[-50:]):.4f}  "
                f"graph captures so far {trainer.captures}")
    # ---- held-out rollout MSE of both learned weight sets, evaluated by the oracle
    ev = MovingDots(C, H, W, A, Rw, seed=99)
    st, rw, dn, ac = ev.get_trajectories(100, 20)
    st, rw = torch.from_numpy(st).to(dev), torch.from_numpy(rw).to(dev)
    dn = torch.from_numpy(dn.astype(np.float32)).to(dev)
    sd_ours = {k: {n: v.detach().clone() for n, v in m.state_dict().items()} for k, m in nets.items()}
    sd_orac = {k: {n: v.detach().clone() for n, v in sd.items()} for k, sd in onets.items()}
    mse_ours = R.measure_prediction_mse(sd_ours, st, rw, dn, ac, num_actions=A)[0]
    mse_orac = R.measure_prediction_mse(sd_orac, st, rw, dn, ac, num_actions=A)[0]
    return {"steps": steps, "batch": batch, "horizon": horizon, "schedule": bool(schedule),
            "theta": "exact (it / steps), device scalar", "graph_captures": trainer.captures,
            "loss_ours": ours, "loss_oracle": oracle,
            "rollout_mse_ours": mse_ours, "rollout_mse_oracle": mse_orac,
            "seconds_ours": t_ours, "seconds_oracle": t_oracle}


def summarize(res, window=50):
    so, sr = smooth(res["loss_ours"], window), smooth(res["loss_oracle"], window)
    rel = np.abs(so - sr) / np.maximum(np.abs(sr), 1e-9)
    mo, mr = np.asarray(res["rollout_mse_ours"]), np.asarray(res["rollout_mse_oracle"])
    return {"smoothed_loss_rel_diff_max": float(rel.max()), "smoothed_loss_rel_diff_mean": float(rel.mean()),
            "final_loss_ours": float(so[-1]), "final_loss_oracle": float(sr[-1]),
            "initial_loss": float(np.mean(res["loss_oracle"][:5])),
            "rollout_mse_mean_ours": float(mo.mean()), "rollout_mse_mean_oracle": float(mr.mean()),
            "rollout_mse_rel_diff_max": float((np.abs(mo - mr) / np.maximum(mr, 1e-9)).max())}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--horizon", type=int, default=5)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "curve_parity.json"))
    ap.add_argument("--rng-seed", type=int, default=None, help="seed of the Bernoulli streams only (same init/data)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--schedule", action="store_true", help="reference horizon schedule 3 .. --horizon (main.py:143-147)")
    args = ap.parse_args()
    res = run(args.steps, args.batch, args.horizon, rng_seed=args.rng_seed, use_graph=not args.no_graph,
              schedule=args.schedule)
    res["summary"] = summarize(res)
    res["summary"]["graph_captures"] = res["graph_captures"]
    print(json.dumps(res["summary"], indent=1))
    print(f"time: ours {res['seconds_ours']:.1f}s, oracle (torch fp32 on the same GPU) {res['seconds_oracle']:.1f}s")
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(res, f)
