"""Single-shape conv launch loop for ncu: python profiles/microbench_one.py CIN N [iters] [gate]
With `gate`, the launch is a dgrad-style conv whose epilogue multiplies by the LeakyReLU derivative of a saved plane
(the planes rotate over more memory than L2 holds, as in the training step)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scm_gan_b200 import kernels as K
cin, n = int(sys.argv[1]), int(sys.argv[2])
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 6
gate = len(sys.argv) > 4 and sys.argv[4] == "gate"
dev = "cuda"; B, H, W = 32, 64, 64
NB = 6
xs = [torch.randn(B, H + 2, W + 2, cin, device=dev).to(torch.bfloat16) for _ in range(NB)]
ys = [K.new_plane(B, H, W, max(n, 16), dev) for _ in range(NB)]
gs = [torch.randn(B, H + 2, W + 2, n, device=dev).to(torch.bfloat16) for _ in range(NB)] if gate else None
w = (torch.randn(9, n, cin, device=dev) * 0.03).to(torch.bfloat16)
bias = torch.zeros(n, device=dev)
def run(i):
    if gate:
        K.conv3x3(xs[i % NB], w, B, H, W, cin=cin, out=ys[i % NB], wrap=True, gate=gs[i % NB], dgrad=True)
    else:
        K.conv3x3(xs[i % NB], w, B, H, W, cin=cin, bias=bias, act=K.ACT_LRELU, out=ys[i % NB], wrap=True)
for i in range(iters):
    run(i)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    with torch.cuda.graph(g, stream=s):
        for i in range(24):
            run(i)
g.replay(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
print(f"cin={cin} n={n} gate={gate}: {e0.elapsed_time(e1) / 24 * 1e3:.1f} us per launch")
