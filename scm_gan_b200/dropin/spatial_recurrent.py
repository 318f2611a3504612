"""CSRN (convolutional spatial recurrent network) with the reference's interface and parameter layout
(reference spatial_recurrent.py:21-119).  Interface-only in the reference: imported by models.py:14, never
instantiated by main.py, so it is off the measured training step.

Four directional sweeps; per line: one bias-free GRU step on the line's pixels (hidden = previous line's state),
the GRU output is the line's context, and the next hidden state is tanh(Conv1d_k3(output)).  The four context maps
are concatenated and mixed by a 1x1 conv.  Reference quirks are preserved on purpose so outputs match:
the right-to-left sweep stores into the *left* context (spatial_recurrent.py:110), so the fourth context map stays
zero and the left-to-right sweep's contexts are overwritten.
"""
import torch
from torch import nn
from torch.nn import functional as F

from scm_gan_b200 import ops as _ops  # noqa: F401  (registers torch.ops.scmgan.*)


def _explode_init(t, channels):
    # deliberately huge N(0, channels) init: "the RNN gradient should explode, not vanish" (spatial_recurrent.py:9-18)
    t.data.normal_(0, channels)


class CSRN(nn.Module):
    DIRECTIONS = ("down", "up", "left", "right")

    def __init__(self, channels):
        super().__init__()
        self.channels = channels
        self.rnn_in = channels
        self.rnn_out = channels
        for d in self.DIRECTIONS:
            setattr(self, "conv_" + d, nn.Conv1d(channels, channels, kernel_size=3, stride=1, padding=1))
        for d in self.DIRECTIONS:
            setattr(self, "rnn_" + d, nn.GRU(channels, channels, bias=False))
        self.conv_combine = nn.Conv2d(channels * 4, channels, kernel_size=1)
        for d in self.DIRECTIONS:
            _explode_init(getattr(self, "conv_" + d).weight, channels)
        for d in self.DIRECTIONS:
            gru = getattr(self, "rnn_" + d)
            _explode_init(gru.weight_hh_l0, channels)
            _explode_init(gru.weight_ih_l0, channels)

    @staticmethod
    def _gru_step(x, h, w_ih, w_hh):
        gi, gh = x @ w_ih.t(), h @ w_hh.t()
        i_r, i_z, i_n = gi.chunk(3, 1)
        h_r, h_z, h_n = gh.chunk(3, 1)
        r, z = torch.sigmoid(i_r + h_r), torch.sigmoid(i_z + h_z)
        n = torch.tanh(i_n + r * h_n)
        return (1 - z) * n + z * h

    def _sweep(self, x, direction, context, order, along_rows):
        b, c, h, w = x.shape
        n = w if along_rows else h
        gru, conv = getattr(self, "rnn_" + direction), getattr(self, "conv_" + direction)
        state = x.new_zeros(b * n, c)
        for i in order:
            line = (x[:, :, i, :] if along_rows else x[:, :, :, i]).permute(0, 2, 1).reshape(b * n, c)
            out = self._gru_step(line, state, gru.weight_ih_l0, gru.weight_hh_l0).view(b, n, c).permute(0, 2, 1)
            if along_rows:
                context[:, :, i, :] = out
            else:
                context[:, :, :, i] = out
            state = torch.tanh(conv(out)).permute(0, 2, 1).reshape(b * n, c)

    def _native_sweep(self, x, direction, along_rows, reverse):
        """One sweep through the hand-written kernels (scmgan::csrn_sweep -> scmgan_gru_conv_sweep_fwd/bwd)."""
        gru, conv = getattr(self, "rnn_" + direction), getattr(self, "conv_" + direction)
        return torch.ops.scmgan.csrn_sweep(x, gru.weight_ih_l0, gru.weight_hh_l0, conv.weight, conv.bias,
                                           along_rows, reverse)[0]

    def forward(self, x):
        b, c, h, w = x.shape
        assert c == self.channels
        if x.is_cuda and 12 * max(h, w) * c * 4 <= 227 * 1024:
            above = self._native_sweep(x, "down", True, False)
            below = self._native_sweep(x, "up", True, True)
            # sic (reference line 110): the right-to-left sweep is stored as the LEFT context, overwriting every
            # column of the left-to-right sweep (whose result is therefore never used), and the right context is zero
            left = self._native_sweep(x, "right", False, True)
            return self.conv_combine(torch.cat((above, below, left, torch.zeros_like(x)), dim=1))
        above, below, left, right = (x.new_zeros(b, c, h, w) for _ in range(4))
        self._sweep(x, "down", above, range(h), True)
        self._sweep(x, "up", below, reversed(range(h)), True)
        self._sweep(x, "left", left, range(w), False)
        self._sweep(x, "right", left, reversed(range(w)), False)  # sic: reference line 110 writes context_left
        return self.conv_combine(torch.cat((above, below, left, right), dim=1))
