"""Throughput of the training step over the configuration grid of SURVEY.md section 8d / BASELINE.json configs 0, 2, 3, 4
(MiniPacMan batch / GPU-count scaling, sc2-shaped frames, rollout-length sweep), several configurations per process so
that the interpreter / NCCL start-up is paid once.  Same timing rules as bench.py (CUDA events around K graph replays
after W warm-up iterations, barrier + synchronize on both sides, max over ranks; inputs device-resident; CF losses every
5th iteration).  One JSON line per configuration (rank 0).

  python profiles/sweep.py --grid minipacman_weak                       # 1 GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
         profiles/sweep.py --grid minipacman_weak,minipacman_strong,sc2,tsweep_short

`scaling`: "weak" = per-GPU batch fixed, "strong" = global batch fixed (split over the ranks).
"""
import argparse
import gc
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (WORKLOADS, algorithmic_flops_per_iter, read_peaks)

CF_RATE = bench.CF_RATE


def grids(world):
    g = {}
    # (workload, per-GPU batch, horizon, cf, scaling, note)
    g["minipacman_weak"] = [("minipacman", 32, 10, True, "weak", "B=32/GPU (reference default batch)")]
    g["minipacman_strong"] = [("minipacman", max(1, bg // world), 10, True, "strong", f"B_global={bg}")
                              for bg in (1024, 4096, 16384) if bg // world <= 4096 and bg % world == 0]
    g["minipacman_batch"] = [("minipacman", b, 10, True, "weak", f"B={b}/GPU") for b in (128, 512, 1024, 2048)]
    g["sc2"] = [("sc2", 256, 10, True, "weak", "B=256/GPU (BASELINE configs[3]: B_global = 256 N)")]
    g["sc2_small"] = [("sc2", 32, 10, True, "weak", "B=32/GPU")]
    g["pong64"] = [("pong64", 32, 10, True, "weak", "B=32/GPU (the bench.py workload)")]
    g["tsweep"] = [("pong64", 32, t + 2, cf, "weak", f"T={t}, CF {'on' if cf else 'off'}")
                   for t in (1, 2, 4, 8, 16, 24, 40) for cf in (False, True)]
    g["tsweep_short"] = [("pong64", 32, t + 2, True, "weak", f"T={t}, CF on") for t in (8, 40)]
    # BASELINE configs[4]: rollout length x counterfactual batch multiplier (counterfactual_horizon k: k - 1 extra
    # Transition calls per CF branch, folded into the rollout batch = up to 3 B samples per call on CF iterations)
    g["tsweep_cf"] = [("pong64", 32, t + 2, True, "weak", f"T={t}, CF on, counterfactual_horizon={k}", k)
                      for t in (1, 4, 8, 16, 40) for k in (1, 2, 3)]
    return g


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", default="minipacman_weak")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    from scm_gan_b200 import synthetic as S
    from scm_gan_b200.train_step import Trainer, build_nets
    peaks = bench.read_peaks()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    todo = []
    for name in args.grid.split(","):
        todo += grids(world)[name]
    for entry in todo:
        workload, B, Hn, cf, scaling, note = entry[:6]
        CF_HORIZON = entry[6] if len(entry) > 6 else bench.CF_HORIZON
        C, H, W, A, Rw = bench.WORKLOADS[workload]
        T = Hn - 2
        nets = build_nets(C, A, Rw, seed=0)
        for n in nets.values():
            n.train()
        nets["transition"]._rng_state[0] += rank
        kw = dict(enable_disentanglement=cf, enable_action_control=cf, counterfactual_horizon=CF_HORIZON)
        trainer = Trainer(nets, loss_kwargs=kw)
        if world > 1:
            from scm_gan_b200.dp import BucketedGradSync
            BucketedGradSync(trainer)
        st, rw, dn, ac = S.synthetic_batch(B, Hn, C, H, W, A, Rw, seed=1234 + rank)
        batch = {"states": st.to(dev), "rewards": rw.to(dev), "dones": dn.to(dev), "actions": torch.as_tensor(ac).to(dev),
                 "cf_indices": torch.randint(16, (B, 2)).to(dev), "cf_perm": torch.randperm(B).to(dev)}

        def run(i):
            return trainer.step(batch, 1.0, cf_now=cf and (i % CF_RATE == 0), use_graph=True)
        for i in range(args.warmup):
            run(i)
        trainer.static_inputs(batch, None, False)
        if cf:
            trainer.static_inputs(batch, None, True)
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            run(i)
        e1.record()
        sync_all()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item() / args.steps
        n_cf = len([i for i in range(args.steps) if cf and i % CF_RATE == 0])
        t_cf = 2 * (CF_HORIZON - 1) * n_cf / args.steps
        flops = bench.algorithmic_flops_per_iter(C, H, W, A, Rw, B, T, t_cf)
        mem = torch.cuda.max_memory_allocated() / 2 ** 30
        if rank == 0:
            print(json.dumps({
                "metric": "training rollout-frames/s", "value": B * world * T / (ms * 1e-3), "unit": "frames/s",
                "n_gpus": world, "ms_per_step": ms, "steps": args.steps, "warmup": args.warmup, "scaling": scaling,
                "config": {"workload": workload, "frame": [C, H, W], "batch_per_gpu": B, "global_batch": B * world,
                           "horizon": Hn, "T": T, "cf": cf, "counterfactual_horizon": CF_HORIZON, "note": note},
                "step_tflops_per_gpu": flops / (ms * 1e-3) / 1e12,
                "frac_of_sustained_bf16_peak": flops / (ms * 1e-3) / 1e12 / peaks["bf16_sustained"],
                "launches_per_step": {f"Hn={k[0]},cf={int(k[1])}": v for k, v in trainer.launches_per_step.items()}, "peak_mem_gib": round(mem, 2)}), flush=True)
        # free everything (graphs hold the pools) before the next configuration
        del trainer, nets, batch
        gc.collect()
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
