"""In-kernel timeline of conv_expand.cuh (SCMGAN_DEBUG bit 4096): clock64 stamps of CTA 0 per warp role and tile.

  SCMGAN_DEBUG=4096 python profiles/expand_timeline.py [--batch 256] [--n 64] [--gated]

Prints, per role and tile iteration, the cycle (relative to the first stamp) at which each phase was reached: that is
how the per-tile costs quoted in conv_expand.cuh / profiles/r02_notes.md were measured.
"""
import argparse
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from scm_gan_b200 import _lib as L  # noqa: E402
from scm_gan_b200 import kernels as K  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--n", type=int, default=64)
ap.add_argument("--gated", action="store_true")
a = ap.parse_args()
B, H, W, n = a.batch, 64, 64, a.n
dt = K.GRAD_DTYPE if a.gated else K.FWD_DTYPE
x = (torch.randn(B, H + 2, W + 2, 16, device="cuda") * 0.5).to(dt)
w = (torch.randn(9, n, 16, device="cuda") * 0.03).to(dt)
out = torch.empty(B, H + 2, W + 2, 128, dtype=dt, device="cuda")
gate = (torch.randn(B, H + 2, W + 2, 128, device="cuda")).to(K.FWD_DTYPE) if a.gated else None
bias = torch.zeros(n, device="cuda")
for _ in range(3):
    K.conv3x3(x, w, B, H, W, cin=16, bias=bias, act=K.ACT_NONE if a.gated else K.ACT_LRELU, out=out, wrap=False, gate=gate,
              dgrad=a.gated)
torch.cuda.synchronize()
buf = (C.c_ulonglong * (3 * 16 * 8))()
fn = L.lib().scmgan_debug_expand_timeline
fn.restype = C.c_int
fn.argtypes = [C.c_void_p, C.c_int]
assert fn(buf, len(buf)) == 0
t = torch.tensor(list(buf), dtype=torch.int64).view(3, 16, 8)
t0 = int(t[t > 0].min())
names = {0: ["loop", "a_empty ok", "copies issued"], 1: ["loop", "acc_empty ok", "a_full ok", "mma issued"],
         2: ["loop", "bias/wait_read", "bar1", "acc_full ok", "tmem+gate", "math+pack", "staged+halo", "bar2"]}
for role, rn in ((0, "producer warp 0"), (1, "mma warp"), (2, "epilogue warp 0")):
    print(rn)
    for it in range(2, 10):
        row = [int(v) - t0 if v > 0 else -1 for v in t[role, it, :len(names[role])]]
        print(f"  tile {it:2d}: " + "  ".join(f"{nm}={v}" for nm, v in zip(names[role], row)))
