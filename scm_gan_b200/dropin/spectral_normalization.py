"""SpectralNorm wrapper with the reference's public surface, backed by the scmgan power-iteration kernels.

Interface contract kept from reference spectral_normalization.py:14-68:
  * `SpectralNorm(module, name='weight', power_iterations=1)`; attributes `.module`, `.name`, `.power_iterations`;
  * the wrapped module loses `<name>` as a parameter and gains `<name>_bar` (trainable), `<name>_u`, `<name>_v`
    (requires_grad=False, unit-normalised N(0,1) draws taken in that order, so a seeded construction reproduces the
    reference's initial state bit for bit and `state_dict()` keys match `convN.module.weight_{bar,u,v}`);
  * every forward call (train *and* eval) runs one power iteration, advances u, v in place and installs
    `<name> = <name>_bar / sigma` before delegating to the wrapped module.
"""
import torch
from torch import nn

from scm_gan_b200 import kernels as K
from scm_gan_b200 import ops  # noqa: F401  (registers the scmgan::* ops)

_EPS = 1e-12


def l2normalize(v, eps=_EPS):
    return v / (v.norm() + eps)


class _NormalisedWeight(torch.autograd.Function):
    """w_bar -> w_bar / sigma, sigma = u.(W v) after the in-place power iteration.  Generic path for wrapped modules
    that are not on the hot path; Encoder and Transition issue the same kernels from their fused ops."""

    @staticmethod
    def forward(ctx, w_bar, u, v):
        sigma = torch.ops.scmgan.spectral_norm_update([w_bar.detach()], [u], [v])
        ctx.save_for_backward(w_bar, sigma)
        ctx.uv = (u, v)  # dereferenced at backward time, as the reference's autograd graph does
        return w_bar.detach() / sigma

    @staticmethod
    def backward(ctx, grad_w):
        w_bar, sigma = ctx.saved_tensors
        u, v = ctx.uv
        grad_bar = torch.empty_like(w_bar)
        scratch = torch.zeros(1, dtype=torch.float32, device=w_bar.device)
        K.spectral_norm_bwd([(grad_w.contiguous().float(), w_bar.detach(), u.detach(), v.detach(), sigma, scratch,
                              grad_bar)])
        return grad_bar, None, None


class SpectralNorm(nn.Module):
    def __init__(self, module, name='weight', power_iterations=1):
        super().__init__()
        self.module = module
        self.name = name
        self.power_iterations = power_iterations
        if not all(hasattr(module, name + s) for s in ("_u", "_v", "_bar")):
            self._install_parameters()

    # -- state ------------------------------------------------------------------------------------------------
    def _install_parameters(self):
        weight = getattr(self.module, self.name)
        rows = weight.shape[0]
        cols = weight.numel() // rows
        # draw order (u then v) and distribution are part of the contract: same seed => same state as the reference
        u = torch.empty(rows, dtype=weight.dtype, device=weight.device).normal_(0, 1)
        v = torch.empty(cols, dtype=weight.dtype, device=weight.device).normal_(0, 1)
        del self.module._parameters[self.name]
        for suffix, value, trainable in (("_u", l2normalize(u), False), ("_v", l2normalize(v), False),
                                         ("_bar", weight.data, True)):
            self.module.register_parameter(self.name + suffix, nn.Parameter(value, requires_grad=trainable))

    def _triple(self):
        m, n = self.module, self.name
        return getattr(m, n + "_bar"), getattr(m, n + "_u"), getattr(m, n + "_v")

    # -- forward ----------------------------------------------------------------------------------------------
    def _update_u_v(self):
        w_bar, u, v = self._triple()
        with torch.no_grad():
            for _ in range(self.power_iterations - 1):  # the reference's models always use 1
                torch.ops.scmgan.spectral_norm_update([w_bar], [u], [v])
        setattr(self.module, self.name, _NormalisedWeight.apply(w_bar, u, v))

    def forward(self, *args):
        self._update_u_v()
        return self.module.forward(*args)
