"""Per-parameter gradient error of the training step as a function of operand precision.

The fp32 oracle (oracle/restated.py, the pinned restatement of reference main.py:155-285) is evaluated once in plain
fp32 and once per precision variant on the SAME weights, inputs and Bernoulli uniforms; a variant rounds the conv
activations / the normalised weights / the gradients entering each conv backward to the given formats (fp32
accumulation everywhere, exactly the contract of a tensor-core path).  For every parameter tensor the relative L2
distance and the cosine against the fp32 gradient are recorded, plus the fraction of LeakyReLU pre-activations that
end up on the other side of the kink (per layer).  Pure oracle arithmetic: none of the product kernels run here, the
script answers "which operand formats can meet the north_star's 1e-2 at all".

  python profiles/grad_precision.py --workload pong64 --batch 32 --horizon 10 --cf-horizon 3 --out profiles/r02_grad_precision.json
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import restated as R  # noqa: E402

WORKLOADS = {"pong64": (3, 64, 64, 4, 1), "minipacman": (3, 15, 19, 5, 2), "sc2": (4, 64, 64, 4, 2)}

VARIANTS = [
    # name, act, w, grad
    ("bf16 act+w+grad (round-1 contract)", torch.bfloat16, torch.bfloat16, torch.bfloat16),
    ("bf16 act+w, fp32 grad", torch.bfloat16, torch.bfloat16, None),
    ("fp32 act+w, bf16 grad", None, None, torch.bfloat16),
    ("bf16 act, fp32 w+grad", torch.bfloat16, None, None),
    ("bf16 w, fp32 act+grad", None, torch.bfloat16, None),
    ("fp16 act+w, bf16 grad", torch.float16, torch.float16, torch.bfloat16),
    ("fp16 act+w, fp32 grad", torch.float16, torch.float16, None),
    ("fp16 act, bf16 w+grad", torch.float16, torch.bfloat16, torch.bfloat16),
    ("tf32 act+w+grad", "tf32", "tf32", "tf32"),
    ("tf32 act+w, bf16 grad", "tf32", "tf32", torch.bfloat16),
]


def fresh(nets0, dev):
    out = {}
    for name, sd in nets0.items():
        d = {}
        for k, v in sd.items():
            t = v.detach().clone().to(dev)
            if t.dtype.is_floating_point and not (k.endswith("_u") or k.endswith("_v") or "bn_conv1" in k):
                t.requires_grad_(True)
            d[k] = t
        out[name] = d
    return out


def run(nets0, batch, A, cf_h, cf_idx, cf_perm, uniforms, dev, act, w, grad):
    nets = fresh(nets0, dev)
    R.ROUND.update(act=act, w=w, grad=grad)
    R.RECORD_SIGNS = []
    try:
        loss, _, z = R.train_step_loss(nets, *batch, num_actions=A, theta=0.5, uniforms=uniforms,
                                       enable_disentanglement=True, enable_action_control=True, cf_now=True,
                                       counterfactual_horizon=cf_h, cf_indices=cf_idx, cf_perm=cf_perm)
        loss.backward()
        signs = R.RECORD_SIGNS
    finally:
        R.ROUND.update(act="inherit", w="inherit", grad="inherit")
        R.RECORD_SIGNS = None
    grads = {f"{n}.{k}": v.grad.detach() for n, sd in nets.items() for k, v in sd.items()
             if v.requires_grad and v.grad is not None}
    return loss.item(), z.detach(), grads, signs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="pong64")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--horizon", type=int, default=10)
    ap.add_argument("--cf-horizon", type=int, default=3)
    ap.add_argument("--out", default=None)
    ap.add_argument("--device", default="cuda" if torch.cuda.is_available() else "cpu")
    args = ap.parse_args()
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dev = args.device
    C, H, W, A, Rw = WORKLOADS[args.workload]
    B, Hn = args.batch, args.horizon
    torch.manual_seed(0)
    nets0 = {"encoder": R.init_encoder(16, C), "decoder": R.init_decoder(16, C),
             "reward_predictor": R.init_reward_predictor(16, Rw), "transition": R.init_transition(16, A)}
    st, rw, dn, ac = R.synthetic_batch(B, Hn, C, H, W, A, Rw, seed=1234, p_done=0.05)
    batch = (st.to(dev), rw.to(dev), dn.to(dev), ac)
    g = torch.Generator().manual_seed(7)
    cf_idx = torch.randint(16, (B, 2), generator=g)
    cf_perm = torch.randperm(B, generator=g)
    gen = torch.Generator(device=dev).manual_seed(3)
    used = []

    def hook(p):  # keep the uniforms 0.02 away from p: the sampled bits then agree across precisions
        u = torch.rand(p.shape, generator=gen, device=p.device)
        u = torch.where((u - p).abs() < 0.02, torch.where(u < p, p - 0.02, p + 0.02), u)
        used.append(u)
        return u

    loss32, z32, g32, s32 = run(nets0, batch, A, args.cf_horizon, cf_idx, cf_perm, hook, dev, None, None, None)
    result = {"workload": args.workload, "shape": [C, H, W], "batch": B, "horizon": Hn, "cf_horizon": args.cf_horizon,
              "loss_fp32": loss32, "device": dev, "variants": []}
    for name, act, w, grad in VARIANTS:
        loss, z, gv, sv = run(nets0, batch, A, args.cf_horizon, cf_idx, cf_perm, list(used), dev, act, w, grad)
        per = {}
        for k, ref in g32.items():
            got = gv[k]
            rel = ((got - ref).norm() / (ref.norm() + 1e-30)).item()
            cos = (torch.dot(got.flatten(), ref.flatten()) / (got.norm() * ref.norm() + 1e-30)).item()
            per[k] = {"rel_l2": rel, "cos": cos}
        flips = {}
        for (tag, a), (_, b) in zip(s32, sv):
            n, f = flips.get(tag, (0, 0))
            flips[tag] = (n + a.numel(), f + int((a != b).sum().item()))
        worst = max(v["rel_l2"] for k, v in per.items() if k.endswith("weight_bar") or k.endswith("weight"))
        entry = {"name": name, "loss": loss, "loss_rel_err": abs(loss - loss32) / abs(loss32),
                 "sample_bit_flips": (z != z32).float().mean().item(),
                 "worst_weight_rel_l2": worst,
                 "median_rel_l2": sorted(v["rel_l2"] for v in per.values())[len(per) // 2],
                 "kink_flip_fraction": {t: f / max(n, 1) for t, (n, f) in flips.items()},
                 "per_parameter": per}
        result["variants"].append(entry)
        print(f"{name:40s} loss_err {entry['loss_rel_err']:.1e}  worst weight rel {worst:.3e}  "
              f"median {entry['median_rel_l2']:.3e}  bit flips {entry['sample_bit_flips']:.1e}", flush=True)
        for k, v in sorted(per.items()):
            print(f"      {k:45s} rel {v['rel_l2']:.3e} cos {v['cos']:.6f}")
        print("      kink flips:", {t: f"{v:.2e}" for t, v in entry["kink_flip_fraction"].items()}, flush=True)
    if args.out:
        with open(args.out, "w") as f:
            json.dump(result, f, indent=1)


if __name__ == "__main__":
    main()
