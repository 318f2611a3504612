"""Thin tensor-level wrappers over the C ABI: they take torch CUDA tensors (memory + stream plumbing only)
and enqueue the hand-written kernels on torch's current stream.  No arithmetic happens in Python/torch here.

A "plane" is a bf16 tensor [B, H+2, W+2, Cs] (NHWC with a one-pixel halo), see include/scmgan.h.
"""
import ctypes as C

import torch

from . import _lib as L
from ._lib import ACT_LRELU, ACT_NONE, ACT_SIGMOID  # noqa: F401

LRELU_SLOPE = 0.01  # F.leaky_relu default used by reference models.py:77-99

# Storage format of the FORWARD operands (activation planes and packed weights).  fp16 (11-bit significand) keeps every
# parameter gradient within 1e-2 of the fp32 reference at the benchmarked shapes; bf16 (the round-1 contract) is
# 3-4x further away, almost all of it from the weight rounding (profiles/r02_grad_precision_*.json).  Gradient planes
# are always bf16: back-propagated values go down to 1e-9, far below fp16's range.
import os as _os
FWD_DTYPE = torch.bfloat16 if _os.environ.get("SCMGAN_FWD_DTYPE", "fp16").lower() in ("bf16", "bfloat16") else torch.float16
GRAD_DTYPE = torch.bfloat16


def _fmt(t):
    """SCMGAN_FMT_* code of a 16-bit tensor."""
    if t.dtype == torch.float16:
        return L.FMT_F16
    assert t.dtype == torch.bfloat16, f"planes / packed weights are bf16 or fp16, got {t.dtype}"
    return L.FMT_BF16


def launch_count():
    """Kernels launched through the C ABI so far (bench.py reports the per-step delta as gpu_launches)."""
    return L.lib().scmgan_launch_count()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def new_plane(B, H, W, Cs, device, dtype=None):
    """dtype: FWD_DTYPE for forward activations, GRAD_DTYPE (default) for gradient planes."""
    # every element (halo included) is written by the producing kernel, so empty() is enough
    return torch.empty((B, H + 2, W + 2, Cs), dtype=dtype or GRAD_DTYPE, device=device)


def fwd_plane(B, H, W, Cs, device):
    return new_plane(B, H, W, Cs, device, FWD_DTYPE)


def pack_nchw(src, dst_plane, c_off=0, c_pad=None, wrap=False, sig=None):
    """fp32 [B,C,H,W] (arbitrary batch stride, dense CHW) -> plane channels [c_off, c_off+c_pad).
    sig: optional contiguous fp32 [B,C,H,W]; the packed value is src * sig * (1 - sig)."""
    B, Cc, H, W = src.shape
    assert src.dtype == torch.float32 and src.is_cuda
    assert src.stride(3) == 1 and src.stride(2) == W and src.stride(1) == H * W, "dense CHW required"
    Cs = dst_plane.shape[3]
    if c_pad is None:
        c_pad = (Cc + 15) // 16 * 16
    assert sig is None or (sig.is_contiguous() and sig.shape == src.shape)
    L.check(L.lib().scmgan_pack_nchw(src.data_ptr(), src.stride(0), Cc, B, H, W, dst_plane.data_ptr(), Cs, c_off,
                                     c_pad, int(wrap), L.ptr(sig), _fmt(dst_plane), _stream()), "scmgan_pack_nchw")


def pack_coords(dst_plane, c_off):
    """CoordConv coordinate channels into plane channels c_off (x) and c_off + 1 (y)."""
    B, Hp, Wp, Cs = dst_plane.shape
    L.check(L.lib().scmgan_pack_coords(dst_plane.data_ptr(), Cs, c_off, B, Hp - 2, Wp - 2, _fmt(dst_plane), _stream()),
            "scmgan_pack_coords")


def coord_wgrad(dy, g, coord_c):
    """Weight gradient of the in-tile coordinate channels coord_c, coord_c + 1: dy [B, Co, H, W] fp32, g [Co, Cin, 3, 3]."""
    B, Co, H, W = dy.shape
    assert dy.is_contiguous() and g.is_contiguous() and g.shape[0] == Co and g.shape[1] >= coord_c + 2
    L.check(L.lib().scmgan_coord_wgrad(dy.data_ptr(), B, Co, H, W, g.data_ptr(), g.stride(0), g.stride(1), coord_c,
                                       _stream()), "scmgan_coord_wgrad")


def pack_weights(jobs):
    """jobs: list of dicts(w, out, sigma, n_pad, k_pad, n_valid, k_valid, s_n, s_k, k_src_off, flip); `out` may be a
    K window (a [:, :, a:b] view) of a wider packed operand."""
    arr = (L.PackJob * len(jobs))()
    for i, j in enumerate(jobs):
        out = j["out"]
        assert out.stride(2) == 1 and out.stride(0) == out.shape[1] * out.stride(1)
        arr[i] = L.PackJob(j["w"].data_ptr(), out.data_ptr(), L.ptr(j.get("sigma")), j["n_pad"], j["k_pad"],
                           j["n_valid"], j["k_valid"], j["s_n"], j["s_k"], j.get("k_src_off", 0), j.get("flip", 0),
                           out.stride(1), _fmt(out))
    L.check(L.lib().scmgan_pack_weights(len(jobs), arr, _stream()), "scmgan_pack_weights")


def packed_weight(n_pad, k_pad, device, dtype=None):
    """GEMM B operand [9][n_pad][k_pad]; FWD_DTYPE unless stated (all weights, forward and dgrad, use it)."""
    return torch.empty((9, n_pad, k_pad), dtype=dtype or FWD_DTYPE, device=device)


def conv3x3(x_plane, w_packed, B, H, W, *, cin, x_c_off=0, scale=1.0, bias=None, sample_bias=None, sample_scale=None,
            act=ACT_NONE,
            out=None, out_c_off=0, wrap=False, add=None, add_c_off=0, gate=None, gate_c_off=0, out_f32=None,
            n_valid=0, sample_out=None, uniforms=None, rng_state=None, dgrad=False, coord_c=None,
            weights_stable=False):
    """weights_stable: the caller guarantees that w_packed was written at least two launches ago in stream order (see
    scmgan_conv_desc::weights_stable); the engine sets it for data-gradient operands, which are packed in the forward."""
    d = _conv_desc(x_plane, w_packed, B, H, W, cin=cin, x_c_off=x_c_off, scale=scale, bias=bias,
                   sample_bias=sample_bias, sample_scale=sample_scale, act=act, out=out, out_c_off=out_c_off, wrap=wrap,
                   add=add, add_c_off=add_c_off, gate=gate, gate_c_off=gate_c_off, out_f32=out_f32, n_valid=n_valid,
                   sample_out=sample_out, uniforms=uniforms, rng_state=rng_state, coord_c=coord_c)
    d.weights_stable = int(bool(weights_stable))
    fn = L.lib().scmgan_conv3x3_dgrad if dgrad else L.lib().scmgan_conv3x3_fwd
    L.check(fn(C.byref(d), _stream()), "scmgan_conv3x3")


def _conv_desc(x_plane, w_packed, B, H, W, *, cin, x_c_off=0, scale=1.0, bias=None, sample_bias=None, sample_scale=None,
               act=ACT_NONE, out=None, out_c_off=0, wrap=False, add=None, add_c_off=0, gate=None, gate_c_off=0,
               out_f32=None, n_valid=0, sample_out=None, uniforms=None, rng_state=None, coord_c=None):
    n = w_packed.shape[1]
    assert w_packed.shape[2] == cin
    d = L.ConvDesc()
    d.B, d.H, d.W = B, H, W
    d.x, d.x_cs, d.x_c_off, d.cin = x_plane.data_ptr(), x_plane.shape[3], x_c_off, cin
    d.w, d.n = w_packed.data_ptr(), n
    d.scale, d.bias, d.sample_bias = scale, L.ptr(bias), L.ptr(sample_bias)
    assert sample_scale is None or (sample_scale.numel() == B and sample_scale.is_contiguous())
    d.sample_scale = L.ptr(sample_scale)
    d.bias_n = 0 if bias is None else min(int(bias.numel()), n)  # shorter than n: the padded channels get no bias
    d.act, d.slope = act, LRELU_SLOPE
    d.out = L.ptr(out)
    d.out_cs = out.shape[3] if out is not None else 0
    d.out_c_off, d.wrap = out_c_off, int(wrap)
    d.add = L.ptr(add)
    d.add_cs = add.shape[3] if add is not None else 0
    d.add_c_off = add_c_off
    d.gate = L.ptr(gate)
    d.gate_cs = gate.shape[3] if gate is not None else 0
    d.gate_c_off = gate_c_off
    d.out_f32, d.n_valid = L.ptr(out_f32), n_valid
    d.sample_out, d.uniforms = L.ptr(sample_out), L.ptr(uniforms)
    d.rng_state = L.ptr(rng_state)  # int64 [2] = {seed, offset}; advanced by the library after the launch
    d.x_fmt, d.w_fmt = _fmt(x_plane), _fmt(w_packed)
    d.out_fmt = _fmt(out) if out is not None else L.FMT_BF16
    d.coord_c1 = 0 if coord_c is None else coord_c + 1   # CoordConv: coordinate channels generated in the tile
    return d


def decoder_bce_fwd(hid_plane, w_packed, T, B, H, W, *, cin, bias, n_valid, dlogits_plane, target_bt, mask_bt, loss_t,
                    logits=None):
    """The decoder's last conv + sigmoid + BCE + masked means in one launch (include/scmgan.h: scmgan_decoder_bce_fwd).
    hid_plane: hidden planes of the T*B decoded latents (t-major); target_bt [B, T, C, H, W] / mask_bt [B, T]: views of
    the batch tensors; loss_t [T] zero-initialised; dlogits_plane: zero-halo gradient plane [T*B, H+2, W+2, 16]."""
    assert target_bt[0, 0].is_contiguous() and target_bt.dtype == torch.float32 and loss_t.numel() == T
    d = L.DecoderBceDesc()
    d.conv = _conv_desc(hid_plane, w_packed, T * B, H, W, cin=cin, bias=bias, act=ACT_NONE, out=dlogits_plane,
                        out_f32=logits, n_valid=n_valid)
    d.target, d.target_bstride, d.target_tstride = target_bt.data_ptr(), target_bt.stride(0), target_bt.stride(1)
    d.mask = L.ptr(mask_bt)
    d.mask_bstride, d.mask_tstride = (mask_bt.stride(0), mask_bt.stride(1)) if mask_bt is not None else (0, 0)
    d.T, d.B = T, B
    d.loss_t = loss_t.data_ptr()
    ws = torch.empty(L.lib().scmgan_decoder_bce_workspace_rows() * T, dtype=torch.float32, device=loss_t.device)
    d.workspace, d.workspace_bytes = ws.data_ptr(), ws.numel() * 4
    L.check(L.lib().scmgan_decoder_bce_fwd(C.byref(d), _stream()), "scmgan_decoder_bce_fwd")


def decoder_bce_bwd(dlogits_plane, g, T, B, H, W):
    """In place: rows of rollout step t of the gradient plane *= g[t] (skipped on the device where g[t] == 1)."""
    assert dlogits_plane.is_contiguous() and g.dtype == torch.float32 and g.numel() == T and g.is_contiguous()
    L.check(L.lib().scmgan_decoder_bce_bwd(dlogits_plane.data_ptr(), dlogits_plane.shape[3], _fmt(dlogits_plane),
                                           g.data_ptr(), T, B, H, W, _stream()), "scmgan_decoder_bce_bwd")


_WGRAD_WS = {}


def _wgrad_workspace(device):
    """Per-device split-K scratch for the weight-gradient kernel (allocated once, reused by every launch: launches
    on one stream are ordered, and the reduce kernel consumes the partials before the next wgrad overwrites them)."""
    key = (device.type, device.index)
    ws = _WGRAD_WS.get(key)
    if ws is None:
        ws = torch.empty(L.lib().scmgan_wgrad_workspace_bytes() // 4, dtype=torch.float32, device=device)
        _WGRAD_WS[key] = ws
    return ws


_WGRAD_RING = {}
_SIDE_STREAM = {}
_RING_BYTES = 320 << 20


class DeferredReduces:
    """The split-K reductions of the weight-gradient launches of ONE module backward, taken off the critical path:
    `wgrad(..., defer=self)` launches only the tensor-core kernel and records the reduction, `kick()` (called by wgrad)
    runs it on a side stream behind an event, and `join()` makes the compute stream wait for all of them.  The
    reduction blocks are small (256 threads, 4-16 KB of shared memory) and co-reside with the next conv kernel's CTAs,
    so the ~9 us per layer they cost disappear behind it.  Every deferred launch gets its own slice of a 320 MB
    scratch ring (partials of different layers are alive at the same time).  Works eagerly and under CUDA-graph
    capture (the fork/join becomes graph edges).  Disabled with SCMGAN_NO_DEFER=1."""
    CAP = 32

    def __init__(self, device):
        import os
        self.enabled = os.environ.get("SCMGAN_NO_DEFER", "0") != "1"
        key = (device.type, device.index)
        if key not in _WGRAD_RING:
            _WGRAD_RING[key] = torch.empty(_RING_BYTES // 4, dtype=torch.float32, device=device)
            _SIDE_STREAM[key] = torch.cuda.Stream(device=device)
        self.ring, self.side = _WGRAD_RING[key], _SIDE_STREAM[key]
        self.jobs = (L.WgradReduceJob * self.CAP)()
        self.count = C.c_int(0)
        self.cursor = C.c_longlong(0)
        self.launched = 0

    def kick(self):
        n_new = self.count.value - self.launched
        if n_new <= 0:
            return
        main = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(main)
        self.side.wait_event(ev)
        first = C.cast(C.byref(self.jobs, self.launched * C.sizeof(L.WgradReduceJob)), C.POINTER(L.WgradReduceJob))
        L.check(L.lib().scmgan_wgrad_reduce(n_new, first, self.side.cuda_stream), "scmgan_wgrad_reduce")
        self.launched = self.count.value

    def side_section(self):
        """Context manager: kernels launched inside run on the side stream, after everything enqueued so far on the
        compute stream (small reductions nothing on the critical path waits for: bias column sums, the folded action
        weight gradient).  Their inputs must stay alive until join()."""
        import contextlib
        if not self.enabled:
            return contextlib.nullcontext()
        main = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(main)
        self.side.wait_event(ev)
        self.forked = True
        return torch.cuda.stream(self.side)

    def join(self):
        if self.launched or getattr(self, "forked", False):
            torch.cuda.current_stream().wait_stream(self.side)


def wgrad(dy_plane, x_plane, g, B, H, W, *, cout, cin, dy_c_off=0, x_c_off=0, g_s_co, g_s_ci, g_s_tap=1, flip=False,
          co_valid=None, ci_valid=None, scale=1.0, db=None, defer=None):
    d = L.WgradDesc()
    d.B, d.H, d.W = B, H, W
    d.dy, d.dy_cs, d.dy_c_off, d.cout = dy_plane.data_ptr(), dy_plane.shape[3], dy_c_off, cout
    d.x, d.x_cs, d.x_c_off, d.cin = x_plane.data_ptr(), x_plane.shape[3], x_c_off, cin
    d.g, d.g_s_co, d.g_s_ci, d.g_s_tap = g.data_ptr(), g_s_co, g_s_ci, g_s_tap
    d.flip = int(flip)
    d.co_valid = cout if co_valid is None else co_valid
    d.ci_valid = cin if ci_valid is None else ci_valid
    d.scale = scale
    d.db = L.ptr(db)  # bias gradient accumulated alongside (zero-initialised by the caller)
    d.dy_fmt, d.x_fmt = _fmt(dy_plane), _fmt(x_plane)
    if defer is not None and defer.enabled:
        d.workspace, d.workspace_bytes = defer.ring.data_ptr(), defer.ring.numel() * 4
        d.defer_jobs, d.defer_cap = defer.jobs, defer.CAP
        d.defer_count, d.workspace_cursor = C.pointer(defer.count), C.pointer(defer.cursor)
        L.check(L.lib().scmgan_conv3x3_wgrad(C.byref(d), _stream()), "scmgan_conv3x3_wgrad")
        defer.kick()
        return
    ws = _wgrad_workspace(g.device)
    d.workspace, d.workspace_bytes = ws.data_ptr(), ws.numel() * 4
    L.check(L.lib().scmgan_conv3x3_wgrad(C.byref(d), _stream()), "scmgan_conv3x3_wgrad")


def plane_colsum(plane, c_off, n, B, H, W, S=None, db=None):
    L.check(L.lib().scmgan_plane_colsum(plane.data_ptr(), plane.shape[3], c_off, n, B, H, W, L.ptr(S), L.ptr(db),
                                        _stream()), "scmgan_plane_colsum")


def spectral_norm_fwd(layers):
    """layers: list of (wbar, u, v, sigma, u_save, v_save); wbar is [rows, ...]."""
    arr = (L.SnLayer * len(layers))()
    for i, (w, u, v, s, us, vs) in enumerate(layers):
        rows = w.shape[0]
        arr[i] = L.SnLayer(w.data_ptr(), u.data_ptr(), v.data_ptr(), s.data_ptr(), L.ptr(us), L.ptr(vs), rows,
                           w.numel() // rows)
    L.check(L.lib().scmgan_spectral_norm_fwd(len(layers), arr, _stream()), "scmgan_spectral_norm_fwd")


def spectral_norm_fwd_n(wbar, u, v, sigma):
    """sigma [iters, n_layers] (contiguous fp32): `iters` successive power iterations of every layer in one launch;
    sigma[i, l] is what call i of layer l's SpectralNorm.forward would compute.  u, v end up advanced `iters` times."""
    iters, n = sigma.shape
    assert sigma.is_contiguous() and n == len(wbar)
    arr = (L.SnLayer * n)()
    for i, w in enumerate(wbar):
        rows = w.shape[0]
        arr[i] = L.SnLayer(w.data_ptr(), u[i].data_ptr(), v[i].data_ptr(), sigma[0, i:i + 1].data_ptr(), None, None, rows,
                           w.numel() // rows)
    L.check(L.lib().scmgan_spectral_norm_fwd_n(n, arr, iters, n, _stream()), "scmgan_spectral_norm_fwd_n")


def spectral_norm_bwd(layers):
    """layers: list of (g, wbar, u, v, sigma, dot, out[, accumulate[, sigma2]])."""
    arr = (L.SnBwdLayer * len(layers))()
    for i, lay in enumerate(layers):
        g, w, u, v, s, dot, out = lay[:7]
        rows = w.shape[0]
        arr[i] = L.SnBwdLayer(g.data_ptr(), w.data_ptr(), u.data_ptr(), v.data_ptr(), s.data_ptr(), dot.data_ptr(),
                              out.data_ptr(), rows, w.numel() // rows, int(bool(lay[7])) if len(lay) > 7 else 0,
                              L.ptr(lay[8]) if len(lay) > 8 else None)
    L.check(L.lib().scmgan_spectral_norm_bwd(len(layers), arr, _stream()), "scmgan_spectral_norm_bwd")


def _csrn_desc(x, along_rows, reverse, w_ih, w_hh, conv_w, conv_b, ctx, states):
    B, Cc, H, W = x.shape
    assert x.is_contiguous() and ctx.is_contiguous() and x.dtype == torch.float32
    d = L.CsrnSweepDesc()
    d.x, d.xs_b, d.xs_c = x.data_ptr(), Cc * H * W, H * W
    if along_rows:   # lines are image rows
        d.xs_line, d.xs_pix, d.L, d.n = W, 1, H, W
    else:            # lines are image columns
        d.xs_line, d.xs_pix, d.L, d.n = 1, W, W, H
    d.B, d.C, d.reverse = B, Cc, int(reverse)
    d.w_ih, d.w_hh = w_ih.data_ptr(), w_hh.data_ptr()
    d.conv_w, d.conv_b = conv_w.data_ptr(), conv_b.data_ptr()
    d.ctx, d.states = ctx.data_ptr(), L.ptr(states)
    return d


def csrn_sweep_fwd(x, along_rows, reverse, w_ih, w_hh, conv_w, conv_b, ctx, states=None):
    """One CSRN sweep (see include/scmgan.h): x, ctx [B,C,H,W] dense fp32; states [B, L, n, C] or None."""
    d = _csrn_desc(x, along_rows, reverse, w_ih, w_hh, conv_w, conv_b, ctx, states)
    L.check(L.lib().scmgan_gru_conv_sweep_fwd(C.byref(d), _stream()), "scmgan_gru_conv_sweep_fwd")


def csrn_sweep_bwd(x, along_rows, reverse, w_ih, w_hh, conv_w, conv_b, ctx, states, dctx, dx, dparams):
    d = _csrn_desc(x, along_rows, reverse, w_ih, w_hh, conv_w, conv_b, ctx, states)
    assert dctx.is_contiguous() and dx.is_contiguous() and dparams.is_contiguous()
    d.dctx, d.dx, d.dparams = dctx.data_ptr(), dx.data_ptr(), dparams.data_ptr()
    L.check(L.lib().scmgan_gru_conv_sweep_bwd(C.byref(d), _stream()), "scmgan_gru_conv_sweep_bwd")


def philox_uniform(out, rng_state):
    """out (fp32, contiguous) <- the next out.numel() uniforms of the device Philox stream rng_state (int64 [2])."""
    L.check(L.lib().scmgan_philox_uniform(out.data_ptr(), out.numel(), rng_state.data_ptr(), _stream()),
            "scmgan_philox_uniform")


def action_bias(wbar, sigma, bias, act, latent, out):
    B, A = act.shape
    cout = wbar.shape[0]
    L.check(L.lib().scmgan_action_bias(wbar.data_ptr(), L.ptr(sigma), L.ptr(bias), act.data_ptr(), B, cout, latent, A,
                                       out.data_ptr(), _stream()), "scmgan_action_bias")


def action_wgrad(S, act, latent, g):
    B, A = act.shape
    cout = g.shape[0]
    L.check(L.lib().scmgan_action_wgrad(S.data_ptr(), act.data_ptr(), B, cout, latent, A, g.data_ptr(), _stream()),
            "scmgan_action_wgrad")


def bce_logits(x, y, mask, loss, dx=None):
    """x: [B,...] dense fp32 logits; y: same shape, dense per sample (batch stride free)."""
    B = x.shape[0]
    per = x.numel() // B
    assert x.is_contiguous() and y.stride(-1) == 1
    L.check(L.lib().scmgan_bce_logits(x.data_ptr(), y.data_ptr(), y.stride(0), L.ptr(mask), B, per, loss.data_ptr(),
                                      L.ptr(dx), _stream()), "scmgan_bce_logits")


def bce_logits_seq(x, y_bt, mask_bt, loss_t, dx=None):
    """x [T*B, ...] dense fp32 logits (t-major); y_bt [B, T, ...] (a view: any batch / time stride, dense per frame);
    mask_bt [B, T] (any strides); loss_t [T] zero-initialised."""
    B, T = y_bt.shape[0], y_bt.shape[1]
    per = x.numel() // (T * B)
    assert x.is_contiguous() and x.shape[0] == T * B and y_bt[0, 0].is_contiguous() and y_bt[0, 0].numel() == per
    L.check(L.lib().scmgan_bce_logits_seq(x.data_ptr(), y_bt.data_ptr(), y_bt.stride(0), y_bt.stride(1),
                                          L.ptr(mask_bt), mask_bt.stride(0), mask_bt.stride(1), T, B, per,
                                          loss_t.data_ptr(), L.ptr(dx), _stream()), "scmgan_bce_logits_seq")


def replay_sample(frames, rewards, actions, ep_len, n_filled, rng_state, states, rewards_out, dones, actions_out,
                  random_start=True, plan=None):
    """Device-side get_trajectories (include/scmgan.h: scmgan_replay_sample).  frames [slots, max_len, ...] f32,
    rewards [slots, max_len, R] f32, actions [slots, max_len] i32, ep_len [slots] i32, n_filled i32 scalar tensor,
    rng_state i64 [2]; outputs states [B, Hn, ...] f32, rewards_out [B, Hn, R], dones [B, Hn], actions_out i64."""
    slots, max_len = frames.shape[0], frames.shape[1]
    B, Hn = states.shape[0], states.shape[1]
    per = frames[0, 0].numel()
    for t in (frames, rewards, actions, ep_len, states, rewards_out, dones, actions_out):
        assert t.is_contiguous()
    assert actions.dtype == torch.int32 and ep_len.dtype == torch.int32 and n_filled.dtype == torch.int32
    assert actions_out.dtype == torch.int64 and states[0, 0].numel() == per
    d = L.ReplayDesc(frames.data_ptr(), rewards.data_ptr(), actions.data_ptr(), ep_len.data_ptr(), n_filled.data_ptr(),
                     slots, max_len, rewards.shape[2], per, B, Hn, int(bool(random_start)), rng_state.data_ptr(),
                     states.data_ptr(), rewards_out.data_ptr(), dones.data_ptr(), actions_out.data_ptr(), L.ptr(plan))
    L.check(L.lib().scmgan_replay_sample(C.byref(d), _stream()), "scmgan_replay_sample")


def eval_sqerr(x, y_bt, out):
    """x [T*B, ...] dense fp32 logits (t-major); y_bt [B, T, ...] view (dense per frame); out [T*B] per-sample MSE of
    sigmoid(x) against y (reference main.py:812-815)."""
    B, T = y_bt.shape[0], y_bt.shape[1]
    per = x.numel() // (T * B)
    assert x.is_contiguous() and x.shape[0] == T * B and y_bt[0, 0].is_contiguous() and y_bt[0, 0].numel() == per
    assert out.is_contiguous() and out.numel() == T * B
    L.check(L.lib().scmgan_eval_sqerr(x.data_ptr(), y_bt.data_ptr(), y_bt.stride(0), y_bt.stride(1), T, B, per,
                                      out.data_ptr(), _stream()), "scmgan_eval_sqerr")


def eval_stats(sqerr, rpred, rewards_bt, dones_bt, table):
    """sqerr [T*B], rpred [T*B, R] (t-major, dense); rewards_bt [B, T, R], dones_bt [B, T] views; table [T, 5]."""
    B, T, R = rewards_bt.shape
    assert sqerr.is_contiguous() and rpred.is_contiguous() and rpred.shape == (T * B, R) and table.is_contiguous()
    assert R == 1 or rewards_bt.stride(2) == 1
    L.check(L.lib().scmgan_eval_stats(sqerr.data_ptr(), rpred.data_ptr(), rewards_bt.data_ptr(), rewards_bt.stride(0),
                                      rewards_bt.stride(1), dones_bt.data_ptr(), dones_bt.stride(0), dones_bt.stride(1),
                                      T, B, R, table.data_ptr(), _stream()), "scmgan_eval_stats")


def masked_mse_seq(pred, target_bt, mask_bt, scale, loss, loss_raw, dpred=None, scale_dev=None):
    """pred [T*B, R] contiguous (t-major); target_bt [B, T, R], mask_bt [B, T] views; loss [1], loss_raw [T]."""
    B, T, R = target_bt.shape
    assert pred.is_contiguous() and pred.shape == (T * B, R) and (R == 1 or target_bt.stride(2) == 1)
    L.check(L.lib().scmgan_masked_mse_seq(pred.data_ptr(), target_bt.data_ptr(), target_bt.stride(0),
                                          target_bt.stride(1), L.ptr(mask_bt), mask_bt.stride(0), mask_bt.stride(1),
                                          T, B, R, float(scale), L.ptr(scale_dev), loss.data_ptr(), L.ptr(loss_raw),
                                          L.ptr(dpred), _stream()), "scmgan_masked_mse_seq")


def masked_mse(pred, target, mask, scale, loss, dpred=None, scale_dev=None, loss_raw=None):
    """pred [B,R] contiguous; target [B,R] with unit inner stride; mask [B] (any stride) or None.
    scale_dev: optional device scalar multiplied into the loss (theta); loss_raw: optional [1] unscaled mean."""
    B, R = pred.shape
    assert pred.is_contiguous() and target.shape == pred.shape and (R == 1 or target.stride(1) == 1)
    L.check(L.lib().scmgan_masked_mse(pred.data_ptr(), target.data_ptr(), target.stride(0), L.ptr(mask),
                                      mask.stride(0) if mask is not None else 0, B, R, float(scale), L.ptr(scale_dev),
                                      loss.data_ptr(), L.ptr(loss_raw), L.ptr(dpred), _stream()), "scmgan_masked_mse")


def reward_head_fwd(y2, R, r, rmap=None):
    B, _, H, W = y2.shape
    L.check(L.lib().scmgan_reward_head_fwd(y2.data_ptr(), B, R, H, W, r.data_ptr(), L.ptr(rmap), _stream()),
            "scmgan_reward_head_fwd")


def reward_head_bwd(y2, dr, R, d2_plane):
    B, _, H, W = y2.shape
    assert d2_plane.shape[3] == 16
    L.check(L.lib().scmgan_reward_head_bwd(y2.data_ptr(), dr.data_ptr(), B, R, H, W, d2_plane.data_ptr(), _stream()),
            "scmgan_reward_head_bwd")


def cf_loss_fwd(za, zb, unswapped, mask, mode, lam, rowmean, loss):
    B, Lz = za.shape[0], za.shape[1]
    HW = za.numel() // (B * Lz)
    L.check(L.lib().scmgan_cf_loss_fwd(za.data_ptr(), zb.data_ptr(), L.ptr(unswapped), mask.data_ptr(), B, Lz, HW, mode,
                                       lam, rowmean.data_ptr(), loss.data_ptr(), _stream()), "scmgan_cf_loss_fwd")


def cf_loss_bwd(za, zb, unswapped, mask, rowmean, gscale, mode, lam, dza, dzb):
    B, Lz = za.shape[0], za.shape[1]
    HW = za.numel() // (B * Lz)
    L.check(L.lib().scmgan_cf_loss_bwd(za.data_ptr(), zb.data_ptr(), L.ptr(unswapped), mask.data_ptr(),
                                       rowmean.data_ptr(), gscale.data_ptr(), B, Lz, HW, mode, lam, L.ptr(dza),
                                       L.ptr(dzb), _stream()), "scmgan_cf_loss_bwd")


def transition_tail(x, uniforms, p, z):
    L.check(L.lib().scmgan_transition_tail(x.data_ptr(), L.ptr(uniforms), x.numel(), L.ptr(p), z.data_ptr(), _stream()),
            "scmgan_transition_tail")


def clip_adam(chunks, lr, beta1, beta2, eps, step, step_dev=None, gscale=1.0):
    """chunks: list of (p, g, m, v, clip[, step_tensor])."""
    arr = (L.AdamChunk * len(chunks))()
    for i, c in enumerate(chunks):
        p, g, m, v, clip = c[:5]
        st = c[5] if len(c) > 5 else None
        arr[i] = L.AdamChunk(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), clip, L.ptr(st))
    L.check(L.lib().scmgan_clip_adam(len(chunks), arr, lr, beta1, beta2, eps, step, L.ptr(step_dev), gscale,
                                     _stream()), "scmgan_clip_adam")
