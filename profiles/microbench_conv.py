"""Micro-benchmark of the conv kernels at the bench workload's shape (B=32, 64x64): CUDA-event timing of single
launches, rotating buffers.  Usage: python profiles/microbench_conv.py [debug_flags]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scm_gan_b200 import kernels as K

dev = "cuda"
B, H, W = 32, 64, 64


def timeit(fn, iters=20):
    """Average kernel time: the launches are captured into a CUDA graph (the Python/ctypes launch path costs ~15 us
    per call and would hide any kernel faster than that) and the graph is replayed between two events."""
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


def planes(c, n=4):
    return [torch.randn(B, H + 2, W + 2, c, device=dev).to(torch.bfloat16) for _ in range(n)]


def conv_case(cin, n, label, **kw):
    xs = planes(cin)
    ys = [K.new_plane(B, H, W, max(n, 16), dev) for _ in range(4)]
    w = (torch.randn(9, n, cin, device=dev) * 0.03).to(torch.bfloat16)
    bias = torch.zeros(n, device=dev)
    us = timeit(lambda i: K.conv3x3(xs[i % 4], w, B, H, W, cin=cin, bias=bias, act=K.ACT_LRELU, out=ys[i % 4], wrap=True, **kw))
    fl = 2.0 * 9 * B * H * W * cin * n
    print(f"conv {label:28s} cin={cin:3d} n={n:3d}: {us:7.1f} us  {fl / us / 1e6:7.1f} TFLOP/s  out {B*(H+2)*(W+2)*n*2/us/1e3:6.1f} GB/s", flush=True)


def wgrad_case(cin, cout, label):
    xs = planes(cin)
    dys = planes(cout)
    g = torch.zeros(cout, cin, 3, 3, device=dev)
    us = timeit(lambda i: K.wgrad(dys[i % 4], xs[i % 4], g, B, H, W, cout=cout, cin=cin, g_s_co=cin * 9, g_s_ci=9))
    fl = 2.0 * 9 * B * H * W * cin * cout
    print(f"wgrad {label:27s} cin={cin:3d} co={cout:3d}: {us:7.1f} us  {fl / us / 1e6:7.1f} TFLOP/s", flush=True)


print("SCMGAN_DEBUG =", os.environ.get("SCMGAN_DEBUG"))
conv_case(128, 128, "128->128 (transition)")
conv_case(256, 128, "256->128 (conv5)")
conv_case(16, 128, "16->128 (conv1 / dgrad6)")
conv_case(256, 16, "256->16 (conv6)")
conv_case(128, 16, "128->16 (dgrad1 / enc conv4)")
conv_case(64, 16, "64->16 (decoder conv2)")
wgrad_case(128, 128, "128x128")
wgrad_case(256, 128, "256x128")
wgrad_case(16, 128, "16x128")
wgrad_case(256, 16, "256x16")
