"""Golden vector for the MPC planner: runs the UNMODIFIED reference `main.compute_rollout_reward` / the action loop of
`main.play` (main.py:356-368, 455-489) on the reference networks (CPU, eval mode = deterministic thresholds) and stores
the per-action scores in tests/golden/planner.pt.  Run in the build container (needs /root/reference):

    python oracle/make_golden_planner.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import make_golden as MG  # noqa: E402
from oracle import shims  # noqa: E402


def load_reference_main():
    """Import the reference's main.py (argparse runs at import: give it an --env) next to its own models module."""
    shims.install_stub_modules()
    names = ("main", "models", "spectral_normalization", "coordconv", "spatial_recurrent", "datasource",
             "causal_graph", "higgins", "utils")
    saved = {n: sys.modules.pop(n) for n in names if n in sys.modules}
    argv, sys.argv = sys.argv, ["main.py", "--env", "none"]
    sys.path.insert(0, shims.REFERENCE_DIR)
    try:
        with shims.cpu_cuda_noop():
            import main as ref_main
    finally:
        sys.argv = argv
        sys.path.remove(shims.REFERENCE_DIR)
        for n in names:
            sys.modules.pop(n, None)
        sys.modules.update(saved)
    return ref_main


def main():
    ref_main = load_reference_main()
    mods = shims.load_reference_modules()
    C, H, W, A, Rw = 3, 15, 19, 5, 2
    nets = MG.build_reference_nets(mods, C, A, Rw, seed=0)
    for n in nets.values():
        n.eval()
    g = torch.Generator().manual_seed(2024)
    z0 = (torch.rand(1, 16, H, W, generator=g) < 0.3).float()
    sn_before = {k: v.clone() for k, v in nets["transition"].state_dict().items()
                 if k.endswith("weight_u") or k.endswith("weight_v")}
    scores = []
    with torch.no_grad(), shims.cpu_cuda_noop():
        for a in range(A):  # main.py:358-364
            z_a = nets["transition"](z0, ref_main.onehot(a, A))
            scores.append(ref_main.compute_rollout_reward(z_a, nets["transition"], nets["reward_predictor"], A, a,
                                                          rollout_depth=12, rollout_policy="noop"))
    scores = torch.stack([s.detach() for s in scores])
    sn_after = {k: v.clone() for k, v in nets["transition"].state_dict().items()
                if k.endswith("weight_u") or k.endswith("weight_v")}
    out = {"config": dict(C=C, H=H, W=W, A=A, R=Rw, seed=0), "z0": z0, "scores": scores,
           "best_action": int(torch.argmax(scores)), "sn_before": sn_before, "sn_after": sn_after}
    path = os.path.join(MG.GOLDEN_DIR, "planner.pt")
    torch.save(out, path)
    print("wrote", path, "scores", [round(float(s), 4) for s in scores], "best", out["best_action"])


if __name__ == "__main__":
    main()
