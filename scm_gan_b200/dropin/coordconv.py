"""CoordConv2d with the reference's interface (reference coordconv.py:5-15): two coordinate channels in [-1, 1)
are appended to the input before an nn.Conv2d built from the caller's arguments (so the caller passes
in_channels + 2).  Interface-only in the reference (imported by models.py:15, never instantiated by main.py).

Coordinate value for column j is -1 + 2j/W and for row i is -1 + 2i/H.  The reference builds the x map as (W, W)
and the y map as (H, H), so only square inputs work there; this implementation accepts any H x W and coincides
with the reference on square inputs.
"""
import torch
from torch import nn


class CoordConv2d(nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()
        self.conv = nn.Conv2d(*args, **kwargs)

    @staticmethod
    def coordinates(height, width, device, dtype):
        xs = -1.0 + 2.0 * torch.arange(width, device=device, dtype=dtype) / width
        ys = -1.0 + 2.0 * torch.arange(height, device=device, dtype=dtype) / height
        return xs.view(1, 1, 1, width).expand(1, 1, height, width), ys.view(1, 1, height, 1).expand(1, 1, height, width)

    def forward(self, x):
        batch_size, _, height, width = x.shape
        cx, cy = self.coordinates(height, width, x.device, x.dtype)
        x = torch.cat([x, cx.expand(batch_size, -1, -1, -1), cy.expand(batch_size, -1, -1, -1)], dim=1)
        return self.conv(x)
