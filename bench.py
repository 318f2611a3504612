#!/usr/bin/env python
"""Benchmark of the scm-gan world-model training step (BASELINE.json metric: training rollout-frames/s).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload pong64|minipacman|sc2]

One "step" = one training iteration of reference main.py:143-296 on one synthetic batch: encoder -> T x {reward
head, decoder + BCE, transition} -> counterfactual losses (every 5th iteration, as CF_REGULARIZATION_RATE) ->
backward -> clip_grad_value_ -> Adam (+ gradient allreduce when N > 1).  frames/s = B_global * T / seconds.
Prints ONE JSON line (rank 0).  `--impl reference` times the reference algorithm's CPU path (the oracle port of
the reference's PyTorch code; the reference itself is not pip-installable and cannot travel to the GPU box).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (C, H, W, A, R)  -- SURVEY.md section 8 shape table
    "pong64": (3, 64, 64, 4, 1),       # BASELINE.json configs[1]: MiniPong / 64x64 frames, full CF regularisation
    "minipacman": (3, 15, 19, 5, 2),   # configs[0] / configs[2]
    "sc2": (4, 64, 64, 4, 2),          # configs[3]
}
CF_RATE = 5      # reference main.py:54
CF_HORIZON = 3   # counterfactual_horizon in {1, 3}; 3 exercises the extra CF transitions


def algorithmic_flops_per_iter(C, H, W, A, R, B, T, t_cf):
    """BASELINE.md section 3 (fwd + dgrad + wgrad, MAC = 2 FLOP, no halo/padding/recompute counted)."""
    f_enc = 18 * H * W * (3 * C * 128 + 2 * 128 * 128 + 128 * 16)
    f_tr = 18 * H * W * ((16 + A) * 128 + 3 * 128 * 128 + 256 * 128 + 256 * 16)
    f_dec = 18 * H * W * (16 * 64 + 64 * 16 * C)
    h2, w2 = (H - 5) // 2 + 1, (W - 5) // 2 + 1
    f_rew = 18 * ((H - 2) * (W - 2) * 16 * 32 + h2 * w2 * 32 * 3 * R)
    return 3 * B * (f_enc + (T + t_cf) * f_tr + T * (f_dec + f_rew))


def read_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture (or None)."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            return json.load(f)["dram_bytes_per_launch"]
    except Exception:
        return None


def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            p = json.load(f)
        return {"bf16_burst": float(p["bf16_tflops"]), "bf16_sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                "hbm": float(p["hbm_gbs"]), "source": "measured"}
    except Exception:
        return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm": 6650.0, "source": "fallback"}


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        if sm:
            # under load = upper half of the samples (the sampler also sees the idle edges of the region)
            out["sm_mhz"] = statistics.median(sorted(sm)[len(sm) // 2:])
            out["sm_max_mhz"] = max(mx)
        out["reasons"] = sorted(reasons)
        out["samples"] = len(sm)
        return out


def cpu_reference_run(workload, horizon, sample_batch, steps, warmup, threads=None):
    """The reference algorithm on host cores: oracle port of models.py + main.py:143-296 (fp32 torch CPU),
    forward + backward + clip + Adam.  Returns (frames/s, ms/step, cores)."""
    import torch
    from oracle import restated as R
    C, H, W, A, Rw = WORKLOADS[workload]
    # all host cores (torchrun exports OMP_NUM_THREADS=1, which would otherwise throttle the baseline)
    if not threads:
        try:
            threads = len(os.sched_getaffinity(0))
        except AttributeError:
            threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    cores = torch.get_num_threads()
    torch.manual_seed(0)
    nets = {"encoder": R.init_encoder(16, C), "decoder": R.init_decoder(16, C),
            "reward_predictor": R.init_reward_predictor(16, Rw), "transition": R.init_transition(16, A)}
    params = []
    for sd in nets.values():
        for k, v in sd.items():
            if v.dtype.is_floating_point and not (k.endswith("_u") or k.endswith("_v") or "bn_conv1" in k):
                v.requires_grad_(True)
                params.append(v)
    opt = torch.optim.Adam(params, lr=1e-4)
    states, rewards, dones, actions = R.synthetic_batch(sample_batch, horizon, C, H, W, A, Rw, seed=1234)
    g = torch.Generator().manual_seed(5)
    times = []
    for it in range(warmup + steps):
        cf_now = (it % CF_RATE == 0)
        cf_idx = torch.randint(16, (sample_batch, 2), generator=g)
        cf_perm = torch.randperm(sample_batch, generator=g)
        t0 = time.perf_counter()
        opt.zero_grad()
        loss, _, _ = R.train_step_loss(nets, states, rewards, dones, actions, num_actions=A, theta=1.0,
                                       enable_disentanglement=True, enable_action_control=True, cf_now=cf_now,
                                       counterfactual_horizon=CF_HORIZON, cf_indices=cf_idx, cf_perm=cf_perm)
        loss.backward()
        torch.nn.utils.clip_grad_value_([p for p in params if p.grad is not None], 0.1)
        opt.step()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    sec = sum(times) / len(times)
    return sample_batch * (horizon - 2) / sec, sec * 1e3, cores


def time_dominant_kernel(B, H, W, iters=30):
    """Average duration of the dominant kernel (3x3 conv 128->128 implicit GEMM, wrap padding, bias+LeakyReLU
    epilogue) at the workload's shape, rotating over > L2 worth of planes.  Launched the way the training step
    launches it - as nodes of a replayed CUDA graph - and timed with CUDA events on the replaying stream."""
    import torch
    from scm_gan_b200 import kernels as K
    dev = "cuda"
    plane_bytes = B * (H + 2) * (W + 2) * 128 * 2
    nbuf = max(2, int(300e6 // (2 * plane_bytes)) + 1)
    # operands in the step's own storage format (kernels.FWD_DTYPE: fp16 activations and weights)
    xs = [torch.randn(B, H + 2, W + 2, 128, device=dev).to(K.FWD_DTYPE) for _ in range(nbuf)]
    ys = [K.fwd_plane(B, H, W, 128, dev) for _ in range(nbuf)]
    w = (torch.randn(9, 128, 128, device=dev) * 0.03).to(K.FWD_DTYPE)
    bias = torch.zeros(128, device=dev)
    for i in range(3):
        K.conv3x3(xs[i % nbuf], w, B, H, W, cin=128, bias=bias, act=K.ACT_LRELU, out=ys[i % nbuf], wrap=True)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            for i in range(iters):
                K.conv3x3(xs[i % nbuf], w, B, H, W, cin=128, bias=bias, act=K.ACT_LRELU, out=ys[i % nbuf], wrap=True)
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    graph.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    flops = 2.0 * 9 * B * H * W * 128 * 128
    return ms, flops


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="pong64", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=32, help="per-GPU batch (reference default, main.py:31)")
    ap.add_argument("--horizon", type=int, default=10, help="prediction horizon Hn (T = Hn - 2 rollout steps)")
    ap.add_argument("--cpu-sample-batch", type=int, default=0,
                    help="batch of the CPU arm's bounded sample (0: --impl reference runs min(batch, 32), i.e. the whole "
                         "default workload; the cpu_baseline leg of the default run uses 8)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--cf-phase", type=int, default=0,
                    help="iteration i applies the CF losses when (i + cf_phase) %% 5 == 0 (profiling aid)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    C, H, W, A, Rw = WORKLOADS[args.workload]
    B, Hn = args.batch, args.horizon
    T = Hn - 2
    warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    config = {"workload": f"{args.workload}: {C}x{H}x{W} frames, A={A}, R={Rw}, B={B}/GPU, horizon {Hn} (T={T}), "
                          f"both CF losses every {CF_RATE}th iteration, counterfactual_horizon={CF_HORIZON}, Adam+clip",
              "global_batch": B * world, "parallelism": f"dp{world}",
              "l2": "per-step working set (saved activations, several GB at 64x64) far exceeds the 126 MB L2"}

    if args.impl == "reference":
        if rank != 0:
            return
        sb = args.cpu_sample_batch or min(B, 32)
        fps, ms, cores = cpu_reference_run(args.workload, Hn, sb, max(args.steps, 1), args.warmup)
        sample = (f"oracle port of the reference step on host CPU (fp32 torch, all host threads), "
                  + (f"the full per-GPU batch of {B}" if sb == B else f"batch {sb} of {B} (per-frame rate)")
                  + f" of the same {C}x{H}x{W} horizon-{Hn} workload")
        if sb != B:
            config["cpu_sample_batch"] = sb
        print(json.dumps({
            "impl": "reference", "metric": "training rollout-frames/s", "value": fps, "unit": "frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config,
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py --impl ours needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL_DEBUG is left as the caller set it (the driver reads the communicator lines); its output goes to
        # stderr so that stdout stays the one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    from scm_gan_b200 import synthetic as R  # synthetic batch generator
    from scm_gan_b200 import kernels as K
    from scm_gan_b200.train_step import Trainer, build_nets

    nets = build_nets(C, A, Rw, seed=0)  # identical weights on every rank
    for n in nets.values():
        n.train()
    nets["transition"]._rng_state[0] += rank  # independent Bernoulli streams per rank
    trainer = Trainer(nets, loss_kwargs=dict(enable_disentanglement=True, enable_action_control=True,
                                             counterfactual_horizon=CF_HORIZON))
    if world > 1:
        from scm_gan_b200.dp import BucketedGradSync
        BucketedGradSync(trainer)
    torch.manual_seed(1234 + rank)
    st, rw, dn, ac = R.synthetic_batch(B, Hn, C, H, W, A, Rw, seed=1234 + rank)
    host = {"states": st.pin_memory(), "rewards": rw.pin_memory(), "dones": dn.pin_memory(),
            "actions": torch.as_tensor(ac).pin_memory(),
            "cf_indices": torch.randint(16, (B, 2)).pin_memory(), "cf_perm": torch.randperm(B).pin_memory()}
    batch = {k: v.to(dev) for k, v in host.items()}
    use_graph = not args.no_graph
    theta = 1.0

    def run_step(i, from_host=False):
        cf_now = ((i + args.cf_phase) % CF_RATE == 0)
        if from_host:
            tgt = trainer.static_inputs(batch, theta, cf_now) if use_graph else batch
            for k, v in host.items():
                tgt[k].copy_(v, non_blocking=True)
            src = tgt
        else:
            src = batch
        return trainer.step(src, theta, cf_now=cf_now, use_graph=use_graph)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for i in range(warmup):
        run_step(i)
    if use_graph:  # make sure both graph variants exist before timing
        trainer.static_inputs(batch, theta, True)
        trainer.static_inputs(batch, theta, False)
    sync_all()

    # ---------------- device-resident timing (value) ----------------
    sampler = ClockSampler(local_rank) if rank == 0 else None
    n0 = K.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    torch.cuda.nvtx.range_push("timed")  # lets `ncu --nvtx --nvtx-include timed/` select exactly this region
    e0.record()
    for i in range(args.steps):
        loss = run_step(i)
    e1.record()
    sync_all()
    torch.cuda.nvtx.range_pop()
    ms_total = e0.elapsed_time(e1)
    # ---------------- end-to-end timing: host buffers in, loss out, every step ----------------
    loss_host = torch.zeros((), dtype=torch.float32).pin_memory()
    sync_all()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # public API for host-resident data: scm_gan_b200.data.InputPipeline (double-buffered pinned-host -> device
    # copies on a copy stream).  Every step's inputs cross PCIe inside the timed region; the copy of step i+1 is in
    # flight while step i computes.
    from scm_gan_b200.data import InputPipeline
    pipe = InputPipeline(trainer, host, depth=2, device=dev)
    if not use_graph:
        pipe = None
    f0.record()
    if pipe is not None:
        pipe.submit(host)
    for i in range(args.steps):
        if pipe is not None:
            if i + 1 < args.steps:
                pipe.submit(host)
            loss = pipe.step(theta, cf_now=((i + args.cf_phase) % CF_RATE == 0), use_graph=True)
        else:
            loss = run_step(i, from_host=True)
        loss_host.copy_(loss.detach(), non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the caller reads the loss every step
    f1.record()
    sync_all()
    ms_e2e = f0.elapsed_time(f1)
    clocks = sampler.stop() if sampler else None

    t = torch.tensor([ms_total, ms_e2e], dtype=torch.float64, device=dev)
    rank_ms = None
    if world > 1:
        gathered = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(gathered, t)
        rank_ms = [round(g[0].item() / args.steps, 4) for g in gathered]   # per-rank device time per step
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e = t.tolist()
    frames = B * world * T * args.steps
    value = frames / (ms_total * 1e-3)
    e2e_value = frames / (ms_e2e * 1e-3)
    h2d = sum(v.numel() * v.element_size() for v in host.values())

    if use_graph:
        per = trainer.launches_per_step
        n_cf = len([i for i in range(args.steps) if (i + args.cf_phase) % CF_RATE == 0])
        launches = per.get((Hn, True), 0) * n_cf + per.get((Hn, False), 0) * (args.steps - n_cf)
    else:
        launches = (K.launch_count() - n0) // 2

    if rank == 0:
        peaks = read_peaks()
        n_cf = len([i for i in range(args.steps) if (i + args.cf_phase) % CF_RATE == 0])
        t_cf_avg = 2 * (CF_HORIZON - 1) * n_cf / args.steps
        flops_iter = algorithmic_flops_per_iter(C, H, W, A, Rw, B, T, t_cf_avg)
        step_tflops = flops_iter * world / (ms_total / args.steps * 1e-3) / 1e12
        k_ms, k_flops = time_dominant_kernel(B, H, W)
        achieved = k_flops / (k_ms * 1e-3) / 1e12
        out = {
            "metric": "training rollout-frames/s", "value": value, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": ("fp16" if K.FWD_DTYPE == torch.float16 else "bf16") + " forward operands, bf16 gradient planes, f32 accumulate",
            "data": "synthetic", "config": config,
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "conv3x3_igemm_v3_kernel<64,9> (CTA-pair tcgen05 conv, 128->128 ch, one Transition conv)",
                         "achieved": achieved, "peak": peaks["bf16_burst"], "unit": "TFLOP/s",
                         "frac": achieved / peaks["bf16_burst"], "traffic": read_traffic(),
                         "peak_source": peaks["source"] + " (burst: kernel timed alone)",
                         "kernel_ms": k_ms, "kernel_flops": k_flops},
            "step_tflops": {"achieved": step_tflops / world, "peak": peaks["bf16_sustained"],
                            "frac": step_tflops / world / peaks["bf16_sustained"],
                            "note": "algorithmic FLOPs/iter (BASELINE.md section 3) per GPU over the whole step"},
            "use_cuda_graph": use_graph,
        }
        if rank_ms is not None:
            out["rank_ms_per_step"] = rank_ms
            if os.environ.get("SCMGAN_DP_NOSYNC") == "1":
                # diagnostic run: gradient exchange switched off, so every rank's time is its own compute time; the
                # spread is the rank skew an exchange has to wait for (NOT a valid training / bench number)
                out["diagnostic"] = "SCMGAN_DP_NOSYNC=1: no gradient exchange, ranks uncoupled; not a bench value"
        if not args.no_cpu_baseline and world == 1:
            sb = args.cpu_sample_batch or 8
            n_it = 8
            fps, ms, cores = cpu_reference_run(args.workload, Hn, sb, n_it, 1)
            out["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                                   "sample": f"oracle port of the reference step (fp32 torch CPU, all host threads), batch "
                                             f"{sb} of the same frame shape/horizon, {n_it} timed iterations "
                                             f"({ms:.0f} ms each) after 1 warm-up"}
        print(json.dumps(out), flush=True)
    if world > 1:
        # captured CUDA graphs keep references into the NCCL communicator; tearing the process group down in that
        # state can block, so synchronise, flush and leave without running the destructors.
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
