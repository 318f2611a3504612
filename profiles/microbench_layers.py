"""Per-layer timing of the convolution kernels as the training step launches them (bench workload: B=32, 64x64; the
stateless heads at B*T = 256), with the step's storage formats (fp16 forward planes, bf16 gradient planes).
Launches are captured into a CUDA graph and replayed between two CUDA events; buffers rotate over > L2.

  python profiles/microbench_layers.py [--batch 32] [--json out.json]
  SCMGAN_NO_EXPAND=1 python profiles/microbench_layers.py     # previous kernels for the 16-input-channel layers
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from scm_gan_b200 import kernels as K  # noqa: E402

dev = "cuda"
HBM = 6546.6  # GB/s, MEASURED_PEAKS.json


def timeit(fn, iters=20):
    for _ in range(3):
        fn(0)
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        fn(0)
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            for i in range(iters):
                fn(i)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--hw", type=int, nargs=2, default=[64, 64])
    ap.add_argument("--json", default=None)
    ap.add_argument("--only", default="", help="substring filter on the layer label")
    args = ap.parse_args()
    B, (H, W) = args.batch, args.hw
    F16, G16 = K.FWD_DTYPE, K.GRAD_DTYPE
    results = []

    def nbuf(bytes_per):
        return max(2, min(8, int(300e6 // max(bytes_per, 1)) + 1))

    def plane(b, c, dt, n):
        return [(torch.randn(b, H + 2, W + 2, c, device=dev) * 0.5).to(dt) for _ in range(n)]

    def case(label, b, cin, n, *, x_dt, out_dt=None, out_cs=None, out_c_off=0, x_cs=None, x_c_off=0, gate_cs=None,
             f32=False, sample_bias=False, act=K.ACT_LRELU, wrap=True, sample=False, dgrad=False):
        if args.only and args.only not in label:
            return
        x_cs = x_cs or cin
        out_cs = out_cs or n
        nb = nbuf(b * (H + 2) * (W + 2) * max(x_cs, out_cs) * 2)
        xs = plane(b, x_cs, x_dt, nb)
        w = (torch.randn(9, n, cin, device=dev) * 0.03).to(x_dt)
        kw = dict(cin=cin, x_c_off=x_c_off, act=act, wrap=wrap, dgrad=dgrad)
        alg = b * H * W * cin * 2  # algorithmic bytes: input once ...
        if f32:
            outs = [torch.empty(b, n, H, W, device=dev) for _ in range(nb)]
            alg += b * H * W * n * 4
            zs = [torch.empty(b, n, H, W, device=dev) for _ in range(nb)] if sample else None
            us_ = [torch.rand(b, n, H, W, device=dev) for _ in range(nb)] if sample else None
            if sample:
                alg += b * H * W * n * 8
        else:
            outs = [torch.empty(b, H + 2, W + 2, out_cs, dtype=out_dt, device=dev) for _ in range(nb)]
            alg += b * H * W * n * 2  # ... output once
        gates = plane(b, gate_cs, F16, nb) if gate_cs else None
        if gate_cs:
            alg += b * H * W * n * 2
        bias = torch.zeros(n, device=dev)
        sb = torch.randn(b, n, device=dev) if sample_bias else None

        def fn(i):
            j = i % nb
            if f32:
                K.conv3x3(xs[j], w, b, H, W, bias=bias, out_f32=outs[j], n_valid=n, sample_out=zs[j] if sample else None,
                          uniforms=us_[j] if sample else None, **kw)
            else:
                K.conv3x3(xs[j], w, b, H, W, bias=None if sb is not None else bias, sample_bias=sb, out=outs[j],
                          out_c_off=out_c_off, gate=gates[j] if gates else None, **kw)
        us = timeit(fn)
        fl = 2.0 * 9 * b * H * W * cin * n
        r = dict(layer=label, batch=b, cin=cin, n=n, us=round(us, 2), tflops=round(fl / us / 1e6, 1),
                 alg_gbs=round(alg / us / 1e3, 1), hbm_frac=round(alg / us / 1e3 / HBM, 3))
        results.append(r)
        print(f"{label:46s} B={b:4d} cin={cin:3d} n={n:3d}: {us:7.1f} us  {r['tflops']:7.1f} TFLOP/s  "
              f"{r['alg_gbs']:7.1f} GB/s algorithmic = {100 * r['hbm_frac']:5.1f} % of HBM", flush=True)

    T = 8
    print("SCMGAN_NO_EXPAND =", os.environ.get("SCMGAN_NO_EXPAND"), "SCMGAN_DEBUG =", os.environ.get("SCMGAN_DEBUG"),
          " FWD_DTYPE =", F16)
    case("tr conv1 16->128 (sample bias, wrap)", B, 16, 128, x_dt=F16, out_dt=F16, out_cs=256, out_c_off=128, sample_bias=True)
    case("tr d pre5 16->128 (gated dgrad, wrap)", B, 16, 128, x_dt=G16, out_dt=G16, out_cs=256, out_c_off=128, x_cs=192,
         x_c_off=128, gate_cs=256, act=K.ACT_NONE, dgrad=True)
    case("tr conv2-4 128->128", B, 128, 128, x_dt=F16, out_dt=F16)
    case("tr conv5 256->128", B, 256, 128, x_dt=F16, out_dt=F16)
    case("tr d pre4/3 128->128 (gated dgrad)", B, 128, 128, x_dt=G16, out_dt=G16, gate_cs=128, act=K.ACT_NONE, dgrad=True)
    case("tr d pre2 256->128 (gated dgrad, K-concat)", B, 256, 128, x_dt=G16, out_dt=G16, gate_cs=256, act=K.ACT_NONE,
         dgrad=True)
    case("tr d pre1 192->128 (gated dgrad, K-concat)", B, 192, 128, x_dt=G16, out_dt=G16, gate_cs=256, act=K.ACT_NONE,
         dgrad=True)
    case("tr conv6 256->16 (sigmoid + sample, f32)", B, 256, 16, x_dt=F16, f32=True, sample=True, act=K.ACT_SIGMOID, wrap=False)
    case("tr dz 128->16 (f32)", B, 128, 16, x_dt=G16, f32=True, act=K.ACT_NONE, wrap=False, dgrad=True)
    case("dec conv1 16->64 (B*T)", B * T, 16, 64, x_dt=F16, out_dt=F16, out_cs=128, wrap=False)
    case("dec conv2 64->16 (f32 logits, B*T)", B * T, 64, 16, x_dt=F16, x_cs=128, f32=True, act=K.ACT_NONE, wrap=False)
    if not args.only or args.only in "dec conv2 + BCE head (fused, B*T)":
        # fused decoder loss head: conv2 + sigmoid + BCE + means in the epilogue, d logits written as a gradient plane
        b, nb = B * T, 3
        hid = plane(b, 128, F16, nb)
        w = (torch.randn(9, 16, 64, device=dev) * 0.03).to(F16)
        bias = torch.zeros(3, device=dev)
        frames = (torch.rand(B, T + 2, 3, H, W, device=dev) < 0.2).float()
        mask = torch.ones(B, T + 2, device=dev)
        d2 = [torch.empty(b, H + 2, W + 2, 16, dtype=G16, device=dev) for _ in range(nb)]
        loss_t = torch.zeros(T, device=dev)

        def fn(i):
            K.decoder_bce_fwd(hid[i % nb], w, T, B, H, W, cin=64, bias=bias, n_valid=3, dlogits_plane=d2[i % nb],
                              target_bt=frames[:, 1:T + 1], mask_bt=mask[:, 1:T + 1], loss_t=loss_t)
        us = timeit(fn)
        alg = b * H * W * (64 * 2 + 16 * 2 + 3 * 4)
        results.append(dict(layer="dec conv2 + BCE head (fused, B*T)", batch=b, cin=64, n=16, us=round(us, 2),
                            alg_gbs=round(alg / us / 1e3, 1), hbm_frac=round(alg / us / 1e3 / HBM, 3)))
        print(f"{'dec conv2 + BCE head (fused, B*T)':46s} B={b:4d} cin= 64 n= 16: {us:7.1f} us  (memset + conv + finalize)  "
              f"{alg / us / 1e3:7.1f} GB/s algorithmic", flush=True)
    case("dec d1 16->64 (gated dgrad, B*T)", B * T, 16, 64, x_dt=G16, out_dt=G16, out_cs=128, gate_cs=128, act=K.ACT_NONE,
         wrap=False, dgrad=True)
    case("dec dz 64->16 (f32, B*T)", B * T, 64, 16, x_dt=G16, x_cs=128, f32=True, act=K.ACT_NONE, wrap=False, dgrad=True)
    case("rew conv1 16->64(32) (B*T)", B * T, 16, 64, x_dt=F16, out_dt=F16, out_cs=128, wrap=False)
    case("rew conv2 32->16 (f32, B*T)", B * T, 32, 16, x_dt=F16, x_cs=128, f32=True, act=K.ACT_NONE, wrap=False)
    case("enc conv1 16->128 (zero pad)", B, 16, 128, x_dt=F16, out_dt=F16, wrap=False)
    # weight gradients (main kernel + reduction, as launched without the side-stream deferral)
    def wcase(label, b, cin, cout, x_cs=None, dy_cs=None):
        if args.only and args.only not in label:
            return
        x_cs, dy_cs = x_cs or max(cin, 16), dy_cs or max(cout, 16)
        nb = nbuf(b * (H + 2) * (W + 2) * max(x_cs, dy_cs) * 2)
        xs = plane(b, x_cs, F16, nb)
        dys = plane(b, dy_cs, G16, nb)
        g = torch.zeros(cout, cin, 3, 3, device=dev)
        db = torch.zeros(max(cout, 16), device=dev)

        def fn(i):
            K.wgrad(dys[i % nb], xs[i % nb], g, b, H, W, cout=dy_cs if cout < 16 else cout, cin=x_cs if cin < 16 else cin,
                    g_s_co=cin * 9, g_s_ci=9, co_valid=cout, ci_valid=cin, db=db)
        us = timeit(fn)
        fl = 2.0 * 9 * b * H * W * cin * cout
        results.append(dict(layer=label, batch=b, cin=cin, n=cout, us=round(us, 2), tflops=round(fl / us / 1e6, 1)))
        print(f"{label:46s} B={b:4d} cin={cin:3d} n={cout:3d}: {us:7.1f} us  {fl / us / 1e6:7.1f} TFLOP/s  (kernel + reduce)",
              flush=True)

    wcase("wgrad 128x128 (conv2-4)", B, 128, 128)
    wcase("wgrad 256->128 (conv5)", B, 256, 128)
    wcase("wgrad 16->128 (conv1, narrow)", B, 16, 128)
    wcase("wgrad 256->16 (conv6, narrow)", B, 256, 16)
    if args.json:
        with open(args.json, "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
