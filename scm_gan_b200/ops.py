"""torch.library custom ops (`scmgan::*`) with registered autograd.

One forward op and one backward op per network; the bodies only sequence hand-written kernels
(scm_gan_b200.engine).  There is no eager/CPU fallback: calling any op with non-CUDA tensors raises.
"""
import weakref
from typing import List, Optional, Sequence

import torch
from torch import Tensor

from . import engine as E
from . import kernels as K


# (u list, v list) of the module issuing the next differentiable forward op; consumed by setup_context.
# The power-iteration vectors are module state read at *backward* time (reference semantics), so they travel
# beside the op instead of through its functional signature.
_UV_SOURCE = []


# Gradient sinks: parameter storage address -> (gradient buffer, callback, parameter).  When a parameter of a differentiable scmgan op has
# a sink, the backward kernels ADD its gradient straight into the buffer (normally the parameter's .grad), autograd
# receives None for it, and `callback(parameter)` runs in place of the post-accumulate-grad hook.  The weights of
# the world model are shared by every unrolled step (reference main.py:162-215), so without sinks autograd launches
# one elementwise add per parameter and step.  Registered by scm_gan_b200.train_step.Trainer; off by default.
_GRAD_SINK = {}


def register_grad_sink(param, grad, callback=None):
    assert grad.shape == param.shape and grad.dtype == torch.float32 and grad.is_contiguous()
    _GRAD_SINK[param.data_ptr()] = (grad, callback, weakref.ref(param))


def clear_grad_sinks():
    _GRAD_SINK.clear()


def _sinks_of(params):
    """-> (present sink buffers, bit mask over `params`)."""
    bufs, mask = [], 0
    for i, p in enumerate(params):
        s = _GRAD_SINK.get(p.data_ptr()) if _GRAD_SINK else None
        if s is not None and s[2]() is None:   # the registering parameter died: its address may have been reused
            del _GRAD_SINK[p.data_ptr()]
            s = None
        # the sink stands in for autograd's own accumulation only while it IS the parameter's .grad: after
        # zero_grad(set_to_none=True), or once another owner has re-homed .grad, autograd gets the gradient back
        if s is not None and s[0].shape == p.shape and p.grad is not None and p.grad.data_ptr() == s[0].data_ptr():
            bufs.append(s[0])
            mask |= 1 << i
    return bufs, mask


def _expand_sinks(sinks, mask, n):
    it = iter(sinks)
    return [next(it) if (mask >> i) & 1 else None for i in range(n)]


def _merge(outs, mask, n):
    """Re-insert None for the sunk slots into the op's compacted output list."""
    it = iter(outs)
    return [None if (mask >> i) & 1 else next(it) for i in range(n)]


def _notify(params, mask):
    for i, p in enumerate(params):
        if (mask >> i) & 1:
            _, cb, ref = _GRAD_SINK[p.data_ptr()]
            if cb is not None:
                cb(ref())


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("scmgan ops run only on CUDA tensors (sm_100a); there is no CPU fallback")


# ------------------------------------------------------------------------------------------------------------
# Transition
# ------------------------------------------------------------------------------------------------------------
@torch.library.custom_op("scmgan::spectral_norm_update", mutates_args=("u", "v"))
def spectral_norm_update(wbar: Sequence[Tensor], u: Sequence[Tensor], v: Sequence[Tensor]) -> Tensor:
    """Power iteration of every SpectralNorm-wrapped conv of one module (reference spectral_normalization.py:28-31);
    u, v advance in place exactly like the reference's `.data` assignments.  Returns sigma per layer."""
    _require_cuda(*wbar)
    return E.spectral_norm_update(list(wbar), list(u), list(v))


# Philox state {seed, offset} of the module issuing the next training-mode forward (device int64[2]).  Like the
# power-iteration vectors it is module state that the kernels advance, so it travels beside the functional op.
_RNG_SOURCE = []


@torch.library.custom_op("scmgan::transition_fwd", mutates_args=())
def transition_fwd(z: Tensor, a: Tensor, wbar: Sequence[Tensor], bias: Sequence[Tensor], sigma: Tensor,
                   w6: Tensor, b6: Tensor, uniforms: Optional[Tensor], training: bool) -> List[Tensor]:
    _require_cuda(z, a, w6)
    rng = _RNG_SOURCE.pop() if _RNG_SOURCE else None
    zn, p, saved = E.transition_forward(z.contiguous().float(), a.contiguous().float(), list(wbar), list(bias),
                                        sigma, w6, b6,
                                        None if uniforms is None else uniforms.contiguous().float(), training, rng)
    return [zn, p] + saved


@torch.library.custom_op("scmgan::transition_bwd", mutates_args=("sinks",))
def transition_bwd(dz_next: Tensor, p: Tensor, a: Tensor, saved: Sequence[Tensor], wbar: Sequence[Tensor],
                   sigma: Tensor, u: Sequence[Tensor], v: Sequence[Tensor], w6: Tensor, sinks: Sequence[Tensor],
                   sink_mask: int) -> List[Tensor]:
    """Returns [dz] + the parameter gradients (dWbar1..5, db1..5, dW6, db6) whose bit in sink_mask is clear; the
    others are added into `sinks`."""
    _require_cuda(dz_next)
    n = len(wbar)
    dz, dwbar, db, dw6, db6 = E.transition_backward(dz_next.contiguous().float(), p, a.contiguous().float(),
                                                    list(saved), list(wbar), sigma, list(u), list(v), w6,
                                                    _expand_sinks(sinks, sink_mask, 2 * n + 2))
    return [dz] + [t for t in dwbar + db + [dw6, db6] if t is not None]


def _transition_setup(ctx, inputs, output):
    ctx.set_materialize_grads(False)  # the saved bf16 planes are outputs too: never build zero grads for them
    z, a, wbar, bias, sigma, w6, b6, uniforms, training = inputs
    ctx.training = training
    ctx.a, ctx.sigma = a, sigma
    ctx.wbar, ctx.w6 = list(wbar), w6
    ctx.bias, ctx.b6 = list(bias), b6
    ctx.nw = len(wbar)
    # u, v are attached by the calling module (ctx.uv_source) and deliberately NOT saved with version tracking:
    # the reference's backward reads whatever u, v the module holds at backward time.
    ctx.uv = _UV_SOURCE.pop()
    ctx.save_for_backward(*output[1:])


def _transition_backward(ctx, grads):
    saved = list(ctx.saved_tensors)
    p, rest = saved[0], saved[1:]
    dz_next = grads[0]
    n = ctx.nw
    if dz_next is None or not ctx.training:
        return (None,) * 9
    u, v = ctx.uv
    params = ctx.wbar + ctx.bias + [ctx.w6, ctx.b6]
    sinks, mask = _sinks_of(params)
    out = torch.ops.scmgan.transition_bwd(dz_next, p, ctx.a, rest, ctx.wbar, ctx.sigma, list(u), list(v), ctx.w6,
                                          sinks, mask)
    dz = out[0]
    g = _merge(out[1:], mask, 2 * n + 2)
    _notify(params, mask)
    return dz, None, g[:n], g[n:2 * n], None, g[2 * n], g[2 * n + 1], None, None


transition_fwd.register_autograd(_transition_backward, setup_context=_transition_setup)


# ------------------------------------------------------------------------------------------------------------
# Encoder
# ------------------------------------------------------------------------------------------------------------
@torch.library.custom_op("scmgan::encoder_fwd", mutates_args=())
def encoder_fwd(x: Tensor, wbar: Sequence[Tensor], bias: Sequence[Tensor], sigma: Tensor, w4: Tensor,
                b4: Tensor) -> List[Tensor]:
    _require_cuda(x, w4)
    B, Cc, H, W = x.shape
    if not (x.dtype == torch.float32 and x.stride(3) == 1 and x.stride(2) == W and x.stride(1) == H * W):
        x = x.contiguous().float()
    z, saved = E.encoder_forward(x, list(wbar), list(bias), sigma, w4, b4)
    return [z] + saved


@torch.library.custom_op("scmgan::encoder_bwd", mutates_args=("sinks",))
def encoder_bwd(dz: Tensor, z: Tensor, saved: Sequence[Tensor], wbar: Sequence[Tensor], sigma: Tensor,
                u: Sequence[Tensor], v: Sequence[Tensor], w4: Tensor, sinks: Sequence[Tensor],
                sink_mask: int) -> List[Tensor]:
    """Returns the parameter gradients (dWbar1..3, db1..3, dW4, db4) whose bit in sink_mask is clear."""
    _require_cuda(dz)
    n = len(wbar)
    dwbar, db, dw4, db4 = E.encoder_backward(dz.contiguous().float(), z, list(saved), list(wbar), sigma, list(u),
                                             list(v), w4, _expand_sinks(sinks, sink_mask, 2 * n + 2))
    return [t for t in dwbar + db + [dw4, db4] if t is not None]


def _encoder_setup(ctx, inputs, output):
    ctx.set_materialize_grads(False)  # the saved bf16 planes are outputs too: never build zero grads for them
    x, wbar, bias, sigma, w4, b4 = inputs
    ctx.wbar, ctx.sigma, ctx.w4 = list(wbar), sigma, w4
    ctx.bias, ctx.b4 = list(bias), b4
    ctx.nw = len(wbar)
    ctx.uv = _UV_SOURCE.pop()
    ctx.save_for_backward(*output)


def _encoder_backward(ctx, grads):
    saved = list(ctx.saved_tensors)
    z, rest = saved[0], saved[1:]
    if grads[0] is None:
        return (None,) * 6
    n = ctx.nw
    u, v = ctx.uv
    params = ctx.wbar + ctx.bias + [ctx.w4, ctx.b4]
    sinks, mask = _sinks_of(params)
    told = set()

    def layer_ready(idxs):  # engine.encoder_backward: these gradients are final in stream order - announce them now
        for i in idxs:
            if (mask >> i) & 1 and i not in told:
                told.add(i)
                _notify([params[i]], 1)
    E.LAYER_READY_HOOK = layer_ready
    try:
        out = torch.ops.scmgan.encoder_bwd(grads[0], z, rest, ctx.wbar, ctx.sigma, list(u), list(v), ctx.w4, sinks,
                                           mask)
    finally:
        E.LAYER_READY_HOOK = None
    g = _merge(out, mask, 2 * n + 2)
    rest_mask = mask
    for i in told:
        rest_mask &= ~(1 << i)
    _notify(params, rest_mask)
    return None, g[:n], g[n:2 * n], None, g[2 * n], g[2 * n + 1]


encoder_fwd.register_autograd(_encoder_backward, setup_context=_encoder_setup)


# ------------------------------------------------------------------------------------------------------------
# Decoder
# ------------------------------------------------------------------------------------------------------------
@torch.library.custom_op("scmgan::decoder_fwd", mutates_args=())
def decoder_fwd(z: Tensor, w1: Tensor, b1: Tensor, w2: Tensor, b2: Tensor) -> List[Tensor]:
    _require_cuda(z, w1)
    logits, saved = E.decoder_forward(z.contiguous().float(), w1, b1, w2.contiguous(), b2.contiguous())
    return [logits] + saved


@torch.library.custom_op("scmgan::decoder_bwd", mutates_args=("sinks",))
def decoder_bwd(dlogits: Tensor, saved: Sequence[Tensor], w1: Tensor, w2: Tensor, sinks: Sequence[Tensor],
                sink_mask: int) -> List[Tensor]:
    """Returns [dz] + the parameter gradients (dW1, db1, dW2, db2) whose bit in sink_mask is clear."""
    _require_cuda(dlogits)
    out = E.decoder_backward(dlogits.contiguous().float(), list(saved), w1, w2.contiguous(),
                             _expand_sinks(sinks, sink_mask, 4))
    return [t for t in out if t is not None]


def _decoder_setup(ctx, inputs, output):
    ctx.set_materialize_grads(False)  # the saved bf16 planes are outputs too: never build zero grads for them
    z, w1, b1, w2, b2 = inputs
    ctx.w1, ctx.w2 = w1, w2
    ctx.params = [w1, b1, w2, b2]
    ctx.save_for_backward(*output[1:])


def _decoder_backward(ctx, grads):
    if grads[0] is None:
        return (None,) * 5
    sinks, mask = _sinks_of(ctx.params)
    out = torch.ops.scmgan.decoder_bwd(grads[0], list(ctx.saved_tensors), ctx.w1, ctx.w2, sinks, mask)
    g = _merge(out[1:], mask, 4)
    _notify(ctx.params, mask)
    return (out[0], *g)


decoder_fwd.register_autograd(_decoder_backward, setup_context=_decoder_setup)


# ------------------------------------------------------------------------------------------------------------
# Decoder + pixel loss of all rollout steps with the loss head in the last convolution's epilogue
# (scmgan_decoder_bce_fwd; used by scm_gan_b200.train_step - main.py keeps decoder() + its own torch loss ops)
# ------------------------------------------------------------------------------------------------------------
@torch.library.custom_op("scmgan::decoder_bce_seq", mutates_args=())
def decoder_bce_seq(z: Tensor, w1: Tensor, b1: Tensor, w2: Tensor, b2: Tensor, target_bt: Tensor,
                    mask_bt: Tensor) -> List[Tensor]:
    """z [T*B, L, H, W] (t-major); w2 / b2 folded over the latent groups; target_bt [B, T, C, H, W], mask_bt [B, T].
    -> [per-step loss terms [T]] + saved planes."""
    _require_cuda(z, w1, target_bt)
    if not (target_bt.dtype == torch.float32 and target_bt[0, 0].is_contiguous()):
        target_bt = target_bt.contiguous().float()
    if mask_bt.dtype != torch.float32:
        mask_bt = mask_bt.float()
    loss_t, saved = E.decoder_bce_forward(z.contiguous().float(), w1, b1, w2.contiguous(), b2.contiguous(), target_bt,
                                          mask_bt)
    return [loss_t] + saved


@torch.library.custom_op("scmgan::decoder_bce_seq_bwd", mutates_args=("sinks", "saved"))
def decoder_bce_seq_bwd(g: Tensor, saved: Sequence[Tensor], w1: Tensor, w2: Tensor, sinks: Sequence[Tensor],
                        sink_mask: int) -> List[Tensor]:
    _require_cuda(g)
    out = E.decoder_bce_backward(g.contiguous().float(), list(saved), w1, w2.contiguous(), g.numel(),
                                 _expand_sinks(sinks, sink_mask, 4))
    return [t for t in out if t is not None]


def _decoder_bce_setup(ctx, inputs, output):
    ctx.set_materialize_grads(False)
    z, w1, b1, w2, b2, _, _ = inputs
    ctx.w1, ctx.w2 = w1, w2
    ctx.params = [w1, b1, w2, b2]
    ctx.save_for_backward(*output[1:])


def _decoder_bce_backward(ctx, grads):
    if grads[0] is None:
        return (None,) * 7
    if getattr(ctx, "bce_consumed", False):
        # the saved gradient plane is rescaled IN PLACE by the upstream gradient (scmgan_decoder_bce_bwd): a second
        # backward through the same forward (retain_graph=True) would apply it twice
        raise RuntimeError("scmgan::decoder_bce_seq supports one backward pass per forward (retain_graph is not "
                           "supported by the fused loss head; use decoder() + scmgan::bce_logits_seq instead)")
    ctx.bce_consumed = True
    sinks, mask = _sinks_of(ctx.params)
    out = torch.ops.scmgan.decoder_bce_seq_bwd(grads[0], list(ctx.saved_tensors), ctx.w1, ctx.w2, sinks, mask)
    g = _merge(out[1:], mask, 4)
    _notify(ctx.params, mask)
    return (out[0], *g, None, None)


decoder_bce_seq.register_autograd(_decoder_bce_backward, setup_context=_decoder_bce_setup)


# ------------------------------------------------------------------------------------------------------------
# Fused sigmoid + BCE + masked mean (used by scm_gan_b200.train_step; main.py keeps its own torch ops)
# ------------------------------------------------------------------------------------------------------------
@torch.library.custom_op("scmgan::bce_logits", mutates_args=())
def bce_logits(logits: Tensor, target: Tensor, mask: Tensor) -> List[Tensor]:
    _require_cuda(logits, target, mask)
    logits = logits.contiguous().float()
    B = logits.shape[0]
    if not (target.dtype == torch.float32 and target[0].is_contiguous()):
        target = target.contiguous().float()
    loss = torch.zeros(1, dtype=torch.float32, device=logits.device)
    dx = torch.empty_like(logits)
    K.bce_logits(logits, target, mask.contiguous().float(), loss, dx)
    return [loss.view(()), dx]


def _bce_setup(ctx, inputs, output):
    ctx.set_materialize_grads(False)  # the saved bf16 planes are outputs too: never build zero grads for them
    ctx.save_for_backward(output[1])


def _bce_backward(ctx, grads):
    (dx,) = ctx.saved_tensors
    g = grads[0]
    return (None if g is None else dx * g), None, None


bce_logits.register_autograd(_bce_backward, setup_context=_bce_setup)


# ------------------------------------------------------------------------------------------------------------
# Sequence forms: the decoder / reward losses of all T rollout steps in one launch (see include/scmgan.h)
# ------------------------------------------------------------------------------------------------------------
@torch.library.custom_op("scmgan::bce_logits_seq", mutates_args=())
def bce_logits_seq(logits: Tensor, target_bt: Tensor, mask_bt: Tensor) -> List[Tensor]:
    """logits [T*B, C, H, W] (t-major); target_bt [B, T, C, H, W] and mask_bt [B, T] are views of the batch tensors.
    -> [per-step loss terms [T], d sum(terms) / d logits]."""
    _require_cuda(logits, target_bt, mask_bt)
    logits = logits.contiguous().float()
    if not (target_bt.dtype == torch.float32 and target_bt[0, 0].is_contiguous()):
        target_bt = target_bt.contiguous().float()
    if mask_bt.dtype != torch.float32:
        mask_bt = mask_bt.float()
    T = target_bt.shape[1]
    terms = torch.zeros(T, dtype=torch.float32, device=logits.device)
    dx = torch.empty_like(logits)
    K.bce_logits_seq(logits, target_bt, mask_bt, terms, dx)
    return [terms, dx]


def _bce_seq_setup(ctx, inputs, output):
    ctx.set_materialize_grads(False)
    ctx.save_for_backward(output[1])


def _bce_seq_backward(ctx, grads):
    (dx,) = ctx.saved_tensors
    g = grads[0]
    if g is None:
        return None, None, None
    T = g.shape[0]
    return (dx.view(T, -1) * g.view(T, 1)).view_as(dx), None, None


bce_logits_seq.register_autograd(_bce_seq_backward, setup_context=_bce_seq_setup)


@torch.library.custom_op("scmgan::masked_mse_seq", mutates_args=())
def masked_mse_seq(pred: Tensor, target_bt: Tensor, mask_bt: Tensor, scale: float,
                   scale_dev: Optional[Tensor]) -> List[Tensor]:
    """pred [T*B, R] (t-major); target_bt [B, T, R], mask_bt [B, T] views.
    -> [scale * scale_dev * sum_t masked mean_t (0-dim), d / d pred, unscaled per-step means [T]]."""
    _require_cuda(pred, target_bt, mask_bt)
    pred = pred.contiguous().float()
    if not (target_bt.dtype == torch.float32 and (target_bt.shape[2] == 1 or target_bt.stride(2) == 1)):
        target_bt = target_bt.contiguous().float()
    if mask_bt.dtype != torch.float32:
        mask_bt = mask_bt.float()
    T = target_bt.shape[1]
    loss = torch.empty(1, dtype=torch.float32, device=pred.device)
    raw = torch.empty(T, dtype=torch.float32, device=pred.device)
    dpred = torch.empty_like(pred)
    K.masked_mse_seq(pred, target_bt, mask_bt, scale, loss, raw, dpred,
                     scale_dev=None if scale_dev is None else scale_dev.reshape(1).float())
    return [loss.view(()), dpred, raw]


masked_mse_seq.register_autograd(lambda ctx, grads: _mse_backward(ctx, grads), setup_context=lambda ctx, inputs, output: _mse_setup(ctx, inputs, output))


# ------------------------------------------------------------------------------------------------------------
# Masked reward MSE (reference main.py:182-186), one kernel for value and gradient
# ------------------------------------------------------------------------------------------------------------
@torch.library.custom_op("scmgan::masked_mse", mutates_args=())
def masked_mse(pred: Tensor, target: Tensor, mask: Tensor, scale: float,
               scale_dev: Optional[Tensor]) -> List[Tensor]:
    """-> [scale * scale_dev * masked mean, d loss / d pred, unscaled masked mean].  scale_dev is a 0-dim / 1-element
    device tensor (theta of main.py:143): no gradient flows to it."""
    _require_cuda(pred, target, mask)
    pred = pred.contiguous().float()
    if not (target.dtype == torch.float32 and (target.shape[1] == 1 or target.stride(1) == 1)):
        target = target.contiguous().float()
    if mask.dtype != torch.float32:
        mask = mask.float()
    loss = torch.empty(1, dtype=torch.float32, device=pred.device)
    raw = torch.empty(1, dtype=torch.float32, device=pred.device)
    dpred = torch.empty_like(pred)
    K.masked_mse(pred, target, mask, scale, loss, dpred,
                 scale_dev=None if scale_dev is None else scale_dev.reshape(1).float(), loss_raw=raw)
    return [loss.view(()), dpred, raw.view(())]


def _mse_setup(ctx, inputs, output):
    ctx.set_materialize_grads(False)
    ctx.save_for_backward(output[1])


def _mse_backward(ctx, grads):
    (dpred,) = ctx.saved_tensors
    g = grads[0]
    return (None if g is None else dpred * g), None, None, None, None


masked_mse.register_autograd(_mse_backward, setup_context=_mse_setup)


# ------------------------------------------------------------------------------------------------------------
# RewardPredictor
# ------------------------------------------------------------------------------------------------------------
@torch.library.custom_op("scmgan::reward_fwd", mutates_args=())
def reward_fwd(z: Tensor, w1: Tensor, b1: Tensor, w2: Tensor, b2: Tensor) -> List[Tensor]:
    _require_cuda(z, w1)
    r, rmap, saved = E.reward_forward(z.contiguous().float(), w1, b1, w2, b2, True)
    return [r, rmap] + saved


@torch.library.custom_op("scmgan::reward_bwd", mutates_args=("sinks",))
def reward_bwd(dr: Tensor, saved: Sequence[Tensor], w1: Tensor, w2: Tensor, sinks: Sequence[Tensor],
               sink_mask: int) -> List[Tensor]:
    """Returns [dz] + the parameter gradients (dW1, db1, dW2, db2) whose bit in sink_mask is clear."""
    _require_cuda(dr)
    out = E.reward_backward(dr.contiguous().float(), list(saved), w1, w2, _expand_sinks(sinks, sink_mask, 4))
    return [t for t in out if t is not None]


def _reward_setup(ctx, inputs, output):
    ctx.set_materialize_grads(False)  # the saved bf16 planes are outputs too: never build zero grads for them
    z, w1, b1, w2, b2 = inputs
    ctx.w1, ctx.w2 = w1, w2
    ctx.params = [w1, b1, w2, b2]
    ctx.save_for_backward(*output[2:])


def _reward_backward(ctx, grads):
    if grads[0] is None:
        return (None,) * 5
    sinks, mask = _sinks_of(ctx.params)
    out = torch.ops.scmgan.reward_bwd(grads[0], list(ctx.saved_tensors), ctx.w1, ctx.w2, sinks, mask)
    g = _merge(out[1:], mask, 4)
    _notify(ctx.params, mask)
    return (out[0], *g)


reward_fwd.register_autograd(_reward_backward, setup_context=_reward_setup)


# ------------------------------------------------------------------------------------------------------------
# Counterfactual regularisers (reference main.py:242-283), one fused reduction kernel each way
# ------------------------------------------------------------------------------------------------------------
@torch.library.custom_op("scmgan::cf_loss", mutates_args=())
def cf_loss(za: Tensor, zb: Tensor, unswapped: Optional[Tensor], mask: Tensor, mode: int, lam: float) -> List[Tensor]:
    _require_cuda(za, zb, mask)
    za, zb = za.contiguous().float(), zb.contiguous().float()
    loss = torch.zeros(1, dtype=torch.float32, device=za.device)
    rowmean = torch.empty((za.shape[0], za.shape[1]), dtype=torch.float32, device=za.device)
    K.cf_loss_fwd(za, zb, None if unswapped is None else unswapped.contiguous().float(), mask.contiguous().float(),
                  mode, lam, rowmean, loss)
    return [loss.view(()), rowmean]


@torch.library.custom_op("scmgan::cf_loss_bwd", mutates_args=())
def cf_loss_bwd(za: Tensor, zb: Tensor, unswapped: Optional[Tensor], mask: Tensor, rowmean: Tensor, gscale: Tensor,
                mode: int, lam: float) -> List[Tensor]:
    _require_cuda(za, zb)
    za, zb = za.contiguous().float(), zb.contiguous().float()
    dza, dzb = torch.empty_like(za), torch.empty_like(zb)
    K.cf_loss_bwd(za, zb, None if unswapped is None else unswapped.contiguous().float(), mask.contiguous().float(),
                  rowmean, gscale.reshape(1).contiguous().float(), mode, lam, dza, dzb)
    return [dza, dzb]


def _cf_setup(ctx, inputs, output):
    ctx.set_materialize_grads(False)
    za, zb, unswapped, mask, mode, lam = inputs
    ctx.mode, ctx.lam = mode, lam
    ctx.unswapped, ctx.mask = unswapped, mask
    ctx.save_for_backward(za, zb, output[1])


def _cf_backward(ctx, grads):
    if grads[0] is None:
        return (None,) * 6
    za, zb, rowmean = ctx.saved_tensors
    dza, dzb = torch.ops.scmgan.cf_loss_bwd(za, zb, ctx.unswapped, ctx.mask, rowmean, grads[0], ctx.mode, ctx.lam)
    return dza, dzb, None, None, None, None


cf_loss.register_autograd(_cf_backward, setup_context=_cf_setup)


# ------------------------------------------------------------------------------------------------------------
# CSRN directional sweep (reference spatial_recurrent.py:61-114; interface-only layer)
# ------------------------------------------------------------------------------------------------------------
@torch.library.custom_op("scmgan::csrn_sweep", mutates_args=())
def csrn_sweep(x: Tensor, w_ih: Tensor, w_hh: Tensor, conv_w: Tensor, conv_b: Tensor, along_rows: bool,
               reverse: bool) -> List[Tensor]:
    """-> [context map (same shape as x), hidden state entering every line [B, L, n, C]]."""
    _require_cuda(x, w_ih, w_hh, conv_w, conv_b)
    x = x.contiguous().float()
    B, Cc, H, W = x.shape
    lines, n = (H, W) if along_rows else (W, H)
    ctx = torch.empty_like(x)
    states = torch.empty((B, lines, n, Cc), dtype=torch.float32, device=x.device)
    K.csrn_sweep_fwd(x, along_rows, reverse, w_ih.contiguous().float(), w_hh.contiguous().float(),
                     conv_w.contiguous().float(), conv_b.contiguous().float(), ctx, states)
    return [ctx, states]


@torch.library.custom_op("scmgan::csrn_sweep_bwd", mutates_args=())
def csrn_sweep_bwd(dctx: Tensor, x: Tensor, ctx: Tensor, states: Tensor, w_ih: Tensor, w_hh: Tensor, conv_w: Tensor,
                   conv_b: Tensor, along_rows: bool, reverse: bool) -> List[Tensor]:
    """-> [dx, dW_ih, dW_hh, dconv_w, dconv_b]."""
    _require_cuda(dctx, x)
    x = x.contiguous().float()
    B, Cc = x.shape[0], x.shape[1]
    dx = torch.empty_like(x)
    per = 9 * Cc * Cc + Cc
    dparams = torch.zeros((B, per), dtype=torch.float32, device=x.device)
    K.csrn_sweep_bwd(x, along_rows, reverse, w_ih.contiguous().float(), w_hh.contiguous().float(),
                     conv_w.contiguous().float(), conv_b.contiguous().float(), ctx, states,
                     dctx.contiguous().float(), dx, dparams)
    tot = dparams.sum(0)  # fixed-order sum over the per-sample partials
    cc3 = 3 * Cc * Cc
    return [dx, tot[:cc3].view(3 * Cc, Cc).clone(), tot[cc3:2 * cc3].view(3 * Cc, Cc).clone(),
            tot[2 * cc3:3 * cc3].view(Cc, Cc, 3).clone(), tot[3 * cc3:].clone()]


def _csrn_setup(ctx, inputs, output):
    ctx.set_materialize_grads(False)
    x, w_ih, w_hh, conv_w, conv_b, along_rows, reverse = inputs
    ctx.flags = (along_rows, reverse)
    ctx.save_for_backward(x, w_ih, w_hh, conv_w, conv_b, output[0], output[1])


def _csrn_backward(ctx, grads):
    if grads[0] is None:
        return (None,) * 7
    x, w_ih, w_hh, conv_w, conv_b, out, states = ctx.saved_tensors
    dx, dwi, dwh, dcw, dcb = torch.ops.scmgan.csrn_sweep_bwd(grads[0], x, out, states, w_ih, w_hh, conv_w, conv_b,
                                                            ctx.flags[0], ctx.flags[1])
    return dx, dwi, dwh, dcw, dcb, None, None


csrn_sweep.register_autograd(_csrn_backward, setup_context=_csrn_setup)


# ------------------------------------------------------------------------------------------------------------
# CoordConv2d (3x3, stride 1, zero padding 1) on the tcgen05 conv kernels (reference coordconv.py; interface only)
# ------------------------------------------------------------------------------------------------------------
@torch.library.custom_op("scmgan::coordconv3x3", mutates_args=())
def coordconv3x3(x: Tensor, w: Tensor, b: Optional[Tensor]) -> List[Tensor]:
    _require_cuda(x, w)
    y, saved = E.coordconv_forward(x.contiguous().float(), w.contiguous().float(),
                                   None if b is None else b.contiguous().float())
    return [y] + saved


@torch.library.custom_op("scmgan::coordconv3x3_bwd", mutates_args=())
def coordconv3x3_bwd(dy: Tensor, saved: Sequence[Tensor], w: Tensor) -> List[Tensor]:
    _require_cuda(dy)
    return list(E.coordconv_backward(dy.contiguous().float(), list(saved), w.contiguous().float()))


def _coordconv_setup(ctx, inputs, output):
    ctx.set_materialize_grads(False)
    x, w, b = inputs
    ctx.w, ctx.has_bias = w, b is not None
    ctx.save_for_backward(*output[1:])


def _coordconv_backward(ctx, grads):
    if grads[0] is None:
        return None, None, None
    dx, dw, db = torch.ops.scmgan.coordconv3x3_bwd(grads[0], list(ctx.saved_tensors), ctx.w)
    return dx, dw, (db if ctx.has_bias else None)


coordconv3x3.register_autograd(_coordconv_backward, setup_context=_coordconv_setup)
