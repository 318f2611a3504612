"""wgrad launch loop for ncu: python profiles/microbench_wgrad_one.py CIN COUT [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scm_gan_b200 import kernels as K
cin, cout = int(sys.argv[1]), int(sys.argv[2])
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 6
dev = "cuda"; B, H, W = 32, 64, 64
xs = [torch.randn(B, H + 2, W + 2, cin, device=dev).to(torch.bfloat16) for _ in range(3)]
dys = [torch.randn(B, H + 2, W + 2, cout, device=dev).to(torch.bfloat16) for _ in range(3)]
g = torch.zeros(cout, cin, 3, 3, device=dev)
for i in range(iters):
    K.wgrad(dys[i % 3], xs[i % 3], g, B, H, W, cout=cout, cin=cin, g_s_co=cin * 9, g_s_ci=9)
torch.cuda.synchronize()
print("done")
