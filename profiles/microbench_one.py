"""Single-shape conv launch loop for ncu: python profiles/microbench_one.py CIN N [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scm_gan_b200 import kernels as K
cin, n = int(sys.argv[1]), int(sys.argv[2])
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 6
dev = "cuda"; B, H, W = 32, 64, 64
xs = [torch.randn(B, H + 2, W + 2, cin, device=dev).to(torch.bfloat16) for _ in range(3)]
ys = [K.new_plane(B, H, W, max(n, 16), dev) for _ in range(3)]
w = (torch.randn(9, n, cin, device=dev) * 0.03).to(torch.bfloat16)
bias = torch.zeros(n, device=dev)
for i in range(iters):
    K.conv3x3(xs[i % 3], w, B, H, W, cin=cin, bias=bias, act=K.ACT_LRELU, out=ys[i % 3], wrap=True)
torch.cuda.synchronize()
print("done")
