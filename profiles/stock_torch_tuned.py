"""TUNED 'stock PyTorch on the same GPU' baseline (cuDNN is the thing
to beat).  The reference algorithm (oracle restatement of main.py:155-296: plain torch ops) on the bench workload with
everything a PyTorch user can switch on without writing kernels:

  eager_fp32_tf32     cudnn.benchmark, TF32 on, eager launches                      (round-1 figure)
  eager_bf16_nhwc     + channels_last activations / weights, torch.autocast(bfloat16)
  graph_bf16_nhwc     + the whole iteration (zero_grad .. clip .. capturable Adam) as ONE CUDA graph per (cf) variant,
                      replayed 4 regular : 1 counterfactual like bench.py

The spectral-norm power iteration writes u, v with copy_() here (the reference rebinds `.data`, which a captured graph
cannot follow); same arithmetic.  Timing: CUDA events over `iters` iterations after warm-up.  One JSON line per variant.

    python profiles/stock_torch_tuned.py [--iters 10] [--workload pong64|sc2|minipacman] [--batch 32]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from oracle import restated as R  # noqa: E402

SHAPES = {"pong64": (3, 64, 64, 4, 1), "sc2": (4, 64, 64, 4, 2), "minipacman": (3, 15, 19, 5, 2)}


def sn_copy(sd, prefix):
    """restated.spectral_norm_weight with in-place u/v updates (capturable)."""
    u, v, w = sd[prefix + "weight_u"], sd[prefix + "weight_v"], sd[prefix + "weight_bar"]
    h = w.shape[0]
    with torch.no_grad():
        w2 = w.view(h, -1).float()
        v.copy_(R.l2normalize(torch.mv(w2.t(), u)))
        u.copy_(R.l2normalize(torch.mv(w2, v)))
    # autograd must not see the buffers the next call overwrites: sigma is taken with copies of this call's u, v (the
    # reference's "last u, v" backward quirk, DESIGN.md section 5, is not reproduced - irrelevant for timing)
    sigma = u.clone().dot(w.view(h, -1).mv(v.clone()))
    return w / sigma.expand_as(w)


_orig_conv2d, _orig_convT2d = torch.nn.functional.conv2d, torch.nn.functional.conv_transpose2d


def _conv2d(x, *a, **k):
    if _conv2d.nhwc:
        x = x.contiguous(memory_format=torch.channels_last)
    return _orig_conv2d(x, *a, **k)


def _convT2d(x, *a, **k):
    if _convT2d.nhwc:
        x = x.contiguous(memory_format=torch.channels_last)
    return _orig_convT2d(x, *a, **k)


_conv2d.nhwc = _convT2d.nhwc = False


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--workload", default="pong64")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--horizon", type=int, default=10)
    ap.add_argument("--variants", default="eager_fp32_tf32,eager_bf16_nhwc,graph_bf16_nhwc")
    args = ap.parse_args()
    dev = "cuda"
    C, H, W, A, Rw = SHAPES[args.workload]
    B, Hn = args.batch, args.horizon
    T = Hn - 2
    torch.backends.cudnn.benchmark = True
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    R.spectral_norm_weight = sn_copy
    R.F.conv2d, R.F.conv_transpose2d = _conv2d, _convT2d

    def pixel_loss_fp32(target, predicted):   # torch refuses F.binary_cross_entropy under autocast: evaluate it in fp32
        with torch.autocast("cuda", enabled=False):
            return _orig_bce(predicted.float(), target.float(), reduction="none").mean(-1).mean(-1).mean(-1)
    _orig_bce = torch.nn.functional.binary_cross_entropy
    R.decoder_pixel_loss = pixel_loss_fp32

    for variant in args.variants.split(","):
        bf16 = "bf16" in variant
        nhwc = "nhwc" in variant
        graphed = variant.startswith("graph")
        torch.manual_seed(0)
        nets = {"encoder": R.init_encoder(16, C), "decoder": R.init_decoder(16, C),
                "reward_predictor": R.init_reward_predictor(16, Rw), "transition": R.init_transition(16, A)}
        params = []
        for sd in nets.values():
            for k in list(sd):
                v = sd[k].to(dev)
                sd[k] = v
                if v.dtype.is_floating_point and not (k.endswith("_u") or k.endswith("_v") or "bn_conv1" in k):
                    v.requires_grad_(True)
                    params.append(v)
        opt = torch.optim.Adam(params, lr=1e-4, capturable=graphed)
        st, rw, dn, ac = R.synthetic_batch(B, Hn, C, H, W, A, Rw, seed=1234)
        st, rw, dn = st.to(dev), rw.to(dev), dn.to(dev)
        ac = torch.as_tensor(ac).to(dev)
        # NHWC: every convolution receives a channels_last input (no-op once activations are channels_last; the
        # parameters stay contiguous because the spectral norm views them as matrices, cuDNN re-lays them per call)
        _conv2d.nhwc = _convT2d.nhwc = nhwc
        g = torch.Generator().manual_seed(1)
        cf_idx, cf_perm = torch.randint(16, (B, 2), generator=g), torch.randperm(B, generator=g).to(dev)

        def iteration(cf_now):
            opt.zero_grad(set_to_none=False)
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=bf16):
                loss, _, _ = R.train_step_loss(nets, st, rw, dn, ac, num_actions=A, theta=1.0,
                                               enable_disentanglement=True, enable_action_control=True, cf_now=cf_now,
                                               counterfactual_horizon=3, cf_indices=cf_idx, cf_perm=cf_perm)
            loss.backward()
            torch.nn.utils.clip_grad_value_([p for p in params if p.grad is not None], 0.1)
            opt.step()
            return loss

        try:
            for p in params:
                p.grad = torch.zeros_like(p)
            run = {}
            if graphed:
                s = torch.cuda.Stream()
                s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s):
                    for cf in (False, True, False):
                        iteration(cf)
                torch.cuda.current_stream().wait_stream(s)
                for cf in (False, True):
                    gr = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(gr):
                        iteration(cf)
                    run[cf] = gr.replay
            else:
                run = {False: lambda: iteration(False), True: lambda: iteration(True)}
                for cf in (False, True, False):
                    run[cf]()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for it in range(args.iters):
                run[(it % 5) == 0]()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.iters
            out = {"variant": variant, "ms_per_step": ms, "frames_per_s": B * T / ms * 1e3}
        except Exception as e:  # noqa: BLE001 - a variant torch cannot run (e.g. an op that is not capturable) is reported, not hidden
            import traceback
            traceback.print_exc()
            out = {"variant": variant, "error": f"{type(e).__name__}: {str(e)[:300]}"}
        out.update({"workload": args.workload, "batch": B, "horizon": Hn, "iters": args.iters,
                    "torch": torch.__version__, "cudnn": torch.backends.cudnn.version(),
                    "device": torch.cuda.get_device_name(0)})
        print(json.dumps(out), flush=True)
        del nets, params, opt
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
