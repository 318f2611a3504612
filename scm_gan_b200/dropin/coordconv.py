"""CoordConv2d with the reference's interface (reference coordconv.py:5-15): two coordinate channels in [-1, 1)
are appended to the input before an nn.Conv2d built from the caller's arguments (so the caller passes
in_channels + 2).  Interface-only in the reference (imported by models.py:15, never instantiated by main.py).

Coordinate value for column j is -1 + 2j/W and for row i is -1 + 2i/H.  The reference builds the x map as (W, W)
and the y map as (H, H), so only square inputs work there; this implementation accepts any H x W and coincides
with the reference on square inputs.
"""
import torch
from torch import nn

from scm_gan_b200 import ops as _ops  # noqa: F401  (registers torch.ops.scmgan.*)


class CoordConv2d(nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()
        self.conv = nn.Conv2d(*args, **kwargs)

    @staticmethod
    def coordinates(height, width, device, dtype):
        xs = -1.0 + 2.0 * torch.arange(width, device=device, dtype=dtype) / width
        ys = -1.0 + 2.0 * torch.arange(height, device=device, dtype=dtype) / height
        return xs.view(1, 1, 1, width).expand(1, 1, height, width), ys.view(1, 1, height, 1).expand(1, 1, height, width)

    def _native_ok(self, x):
        c = self.conv
        return (x.is_cuda and c.kernel_size == (3, 3) and c.stride == (1, 1) and c.padding == (1, 1)
                and c.dilation == (1, 1) and c.groups == 1 and c.padding_mode == "zeros" and x.shape[1] % 2 == 0
                and c.out_channels <= 128)

    def forward(self, x):
        if self._native_ok(x):
            # 3x3 / stride 1 / pad 1: hand-written path - tcgen05 implicit-GEMM conv (16-bit operands, fp32
            # accumulation), dgrad/wgrad kernels behind autograd.  Up to 14 data channels the two coordinate channels
            # are generated inside the conv's im2col tile and never stored (scm_gan_b200/engine.py: coords_in_tile);
            # wider inputs get them written into the 16-bit input plane
            return torch.ops.scmgan.coordconv3x3(x, self.conv.weight, self.conv.bias)[0]
        batch_size, _, height, width = x.shape
        cx, cy = self.coordinates(height, width, x.device, x.dtype)
        x = torch.cat([x, cx.expand(batch_size, -1, -1, -1), cy.expand(batch_size, -1, -1, -1)], dim=1)
        return self.conv(x)
