"""Throughput of the graphed inference consumers against their step-by-step forms:

  eval     measure_prediction_mse at the reference's scale, batch 100 x 100 timesteps (main.py:784-836):
           scm_gan_b200.evaluate.RolloutEvaluator (one CUDA graph, one D2H) vs evaluate.measure_prediction_mse
           (module calls, one D2H) - frames/s = B * (T - 2) decoded frames per evaluation
  planner  one MPC decision of play() (main.py:356-368, 389-391): planner.GraphedPlanner.decide vs the sequential
           reference call order (planner.choose_action) and the folded eager form - decisions/s

    python profiles/inference_bench.py [--json gpurun_out/r02_inference.jsonl]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from scm_gan_b200 import planner, synthetic  # noqa: E402
from scm_gan_b200.evaluate import RolloutEvaluator, measure_prediction_mse  # noqa: E402
from scm_gan_b200.train_step import build_nets  # noqa: E402

SHAPES = {"minipacman": (3, 15, 19, 5, 2), "pong64": (3, 64, 64, 4, 1), "sc2": (4, 64, 64, 4, 2)}
dev = "cuda"


def wall(fn, n):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    ap.add_argument("--workloads", default="minipacman,pong64")
    args = ap.parse_args()
    out = []
    for wl in args.workloads.split(","):
        C, H, W, A, Rw = SHAPES[wl]
        B, T = 100, 100
        nets = build_nets(C, A, Rw, seed=0)
        st, rw, dn, ac = synthetic.synthetic_batch(B, T, C, H, W, A, Rw, seed=3)
        st, rw, dn, ac = st.to(dev), rw.to(dev), dn.to(dev), torch.as_tensor(ac).to(dev)
        ev = RolloutEvaluator(nets)
        ev(st, rw, dn, ac)
        t_graph = wall(lambda: ev(st, rw, dn, ac), 5)
        t_step = wall(lambda: measure_prediction_mse(nets, st, rw, dn, ac), 2)
        out.append({"what": "eval", "workload": wl, "batch": B, "timesteps": T, "graph_ms": t_graph * 1e3,
                    "stepwise_ms": t_step * 1e3, "frames_per_s_graph": B * (T - 2) / t_graph,
                    "frames_per_s_stepwise": B * (T - 2) / t_step, "kernels_per_eval": list(ev.launches.values())[0],
                    "d2h_per_eval": 1})
        print(json.dumps(out[-1]), flush=True)
        for m in nets.values():
            m.eval()
        frames = (torch.rand(1, 3, C, H, W, device=dev) < 0.2).float()
        gp = planner.GraphedPlanner(nets, A)
        gp.decide(frames, 0)
        t_g = wall(lambda: gp.decide(frames, 0), 10)

        def eager(fold):
            with torch.no_grad():
                z = nets["transition"](nets["encoder"](frames), planner.onehot(0, A, dev))
            planner.choose_action(z, nets["transition"], nets["reward_predictor"], A, fold_actions=fold)
        t_seq = wall(lambda: eager(False), 3)
        t_fold = wall(lambda: eager(True), 3)
        out.append({"what": "planner", "workload": wl, "beam": A ** 3, "rollout_depth": 12, "graph_ms": t_g * 1e3,
                    "folded_eager_ms": t_fold * 1e3, "reference_order_ms": t_seq * 1e3,
                    "decisions_per_s_graph": 1 / t_g, "kernels_per_decision": gp.launches})
        print(json.dumps(out[-1]), flush=True)
        del nets, ev, gp
        torch.cuda.empty_cache()
    if args.json:
        with open(args.json, "w") as f:
            for o in out:
                f.write(json.dumps(o) + "\n")


if __name__ == "__main__":
    main()
