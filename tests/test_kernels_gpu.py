"""Kernel-level parity (-m gpu): every C-ABI kernel against a plain fp32 torch statement of the same op.

The tensor-core kernels take 16-bit operands (fp16 for forward activations and weights, bf16 for gradient planes; the
round-1 all-bf16 contract stays selectable); the references below are fed the *same rounded operands* in fp32 (TF32 off),
so the only differences are fp32 summation order and the rounding of stored outputs.  Tolerances are written next to
each check.
"""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _setup():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


def bf(x):
    return x.to(torch.bfloat16).to(torch.float32)


# forward-operand formats under test: fp16 (default product path) and bf16 (round-1 contract, SCMGAN_FWD_DTYPE=bf16)
FWD_DTYPES = [torch.float16, torch.bfloat16]
OUT_TOL = {torch.float16: 6e-4, torch.bfloat16: 4e-3}  # stored-output rounding: 2^-12 / 2^-9 relative per element


def rnd(x, dtype):
    return x.to(dtype).to(torch.float32)


def plane_interior(plane, c_off, C):
    """[B,Hp,Wp,Cs] bf16 -> [B,C,H,W] fp32"""
    return plane[:, 1:-1, 1:-1, c_off:c_off + C].permute(0, 3, 1, 2).float().contiguous()


def ref_plane(x_nchw, wrap):
    """[B,C,H,W] -> padded [B,C,H+2,W+2]"""
    return F.pad(x_nchw, (1, 1, 1, 1), mode="circular" if wrap else "constant")


def ref_conv(x_nchw, w, wrap):
    return F.conv2d(ref_plane(x_nchw, wrap), w)


def make_plane(x_nchw, Cs, c_off, wrap, dtype=torch.bfloat16):
    from scm_gan_b200 import kernels as K
    B, C, H, W = x_nchw.shape
    plane = torch.full((B, H + 2, W + 2, Cs), float("nan"), dtype=dtype, device=DEV)
    K.pack_nchw(x_nchw, plane, c_off=c_off, c_pad=(C + 15) // 16 * 16, wrap=wrap)
    return plane


def pack_conv_weight(w, n_pad, k_pad, sigma=None, dgrad=False, k_src_off=0, k_valid=None, dtype=torch.bfloat16):
    """w: Conv2d weight [Co,Ci,3,3] -> packed [9][n_pad][k_pad]"""
    from scm_gan_b200 import kernels as K
    Co, Ci = w.shape[:2]
    out = K.packed_weight(n_pad, k_pad, w.device, dtype)
    if not dgrad:
        job = dict(w=w, out=out, sigma=sigma, n_pad=n_pad, k_pad=k_pad, n_valid=Co,
                   k_valid=Ci if k_valid is None else k_valid, s_n=Ci * 9, s_k=9, k_src_off=k_src_off, flip=0)
    else:
        job = dict(w=w, out=out, sigma=sigma, n_pad=n_pad, k_pad=k_pad, n_valid=Ci, k_valid=Co, s_n=9, s_k=Ci * 9,
                   flip=1)
    K.pack_weights([job])
    return out


def relerr(a, b):
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def report(name, got, ref, tol):
    err = (got - ref).abs().max().item()
    rel = relerr(got, ref)
    ok = rel <= tol and bool(torch.isfinite(got).all())
    print(f"[{name}] max_abs={err:.3e} rel_l2={rel:.3e} tol={tol:.1e} ref_norm={ref.norm().item():.3e} "
          f"{'OK' if ok else 'FAIL'}", flush=True)
    if not ok:
        d = (got - ref).abs().flatten()
        idx = d.argsort(descending=True)[:5]
        for i in idx.tolist():
            print("    worst idx", i, "got", got.flatten()[i].item(), "ref", ref.flatten()[i].item())
    return ok


# ------------------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("dt", FWD_DTYPES)
@pytest.mark.parametrize("wrap", [False, True])
def test_pack_nchw(wrap, dt):
    _setup()
    torch.manual_seed(1)
    B, C, H, W = 3, 9, 5, 7
    x = torch.randn(B, C, H, W, device=DEV)
    x[0, 0, 0, 0] = 1e5  # beyond fp16's range: saturates to 65504 instead of becoming inf
    plane = make_plane(x, 32, 16, wrap, dt)
    got = plane[:, :, :, 16:32].permute(0, 3, 1, 2).float()
    ref = torch.zeros(B, 16, H + 2, W + 2, device=DEV)
    ref[:, :C] = rnd(ref_plane(x, wrap).clamp(-65504, 65504) if dt == torch.float16 else ref_plane(x, wrap), dt)
    assert torch.equal(got, ref)


CONV_CASES = [
    # (B, H, W, Cin, Cout, wrap, act)
    (3, 15, 19, 128, 128, False, 1),
    (2, 15, 19, 128, 128, True, 1),
    (2, 8, 8, 16, 128, True, 1),
    (2, 6, 10, 64, 64, False, 0),
    (1, 64, 64, 256, 128, True, 1),
    (5, 9, 7, 256, 16, True, 0),
    (2, 16, 16, 128, 48, False, 0),
    (2, 9, 11, 32, 16, False, 0),      # two 16-channel chunks (reward head)
    (2, 64, 64, 256, 128, True, 1),    # Cin = 256: CTA-pair kernel with streamed weights
    (2, 64, 64, 64, 128, False, 1),
    # 16 input channels: image-aligned tiles + TMA-store epilogue (conv_expand.cuh)
    (3, 64, 64, 16, 128, True, 1),
    (2, 64, 64, 16, 64, False, 0),
    (5, 15, 19, 16, 128, True, 1),     # k = 6 rows per tile, last tile of every image clipped (15 = 6 + 6 + 3)
    (3, 15, 19, 16, 64, False, 1),
    (2, 5, 128, 16, 128, True, 0),     # one image row per tile
    (40, 16, 16, 16, 128, False, 1),   # more tiles than CTAs x pipeline depth
]


@pytest.mark.parametrize("dt", FWD_DTYPES)
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_fwd_plane(case, dt):
    """bias + activation epilogue, 16-bit plane output incl. halo handling."""
    _setup()
    from scm_gan_b200 import kernels as K
    B, H, W, Ci, Co, wrap, act = case
    torch.manual_seed(2)
    x = rnd(torch.randn(B, Ci, H, W, device=DEV), dt)
    w = rnd(torch.randn(Co, Ci, 3, 3, device=DEV) / (3.0 * Ci ** 0.5), dt)
    bias = torch.randn(Co, device=DEV)
    xp = make_plane(x, Ci, 0, wrap, dt)
    wp = pack_conv_weight(w, Co, Ci, dtype=dt)
    out = torch.full((B, H + 2, W + 2, Co + 16), float("nan"), dtype=dt, device=DEV)
    K.conv3x3(xp, wp, B, H, W, cin=Ci, bias=bias, act=act, out=out, out_c_off=16, wrap=wrap)
    torch.cuda.synchronize()
    ref = ref_conv(x, w, wrap) + bias.view(1, -1, 1, 1)
    if act == 1:
        ref = F.leaky_relu(ref)
    got = plane_interior(out, 16, Co)
    assert report(f"conv_fwd {case} {dt}", got, ref, OUT_TOL[dt])
    # halo of the produced plane: wrapped copy or zeros
    full = out[:, :, :, 16:16 + Co].permute(0, 3, 1, 2).float()
    assert torch.equal(full, ref_plane(got, wrap)), "halo mismatch"
    # channels outside the written window stay untouched (NaN sentinel)
    assert torch.isnan(out[:, :, :, :16].float()).all()


@pytest.mark.parametrize("dt", FWD_DTYPES)
def test_conv_sample_bias_and_f32_head(dt):
    """Transition conv1-style per-sample bias; conv6-style sigmoid + Bernoulli head with fp32 NCHW outputs."""
    _setup()
    from scm_gan_b200 import kernels as K
    torch.manual_seed(3)
    B, H, W, Ci, Co = 4, 15, 19, 256, 16
    x = rnd(torch.randn(B, Ci, H, W, device=DEV), dt)
    w = rnd(torch.randn(Co, Ci, 3, 3, device=DEV) / (3.0 * Ci ** 0.5), dt)
    bias = torch.randn(Co, device=DEV)
    sb = torch.randn(B, Co, device=DEV)
    u = torch.rand(B, Co, H, W, device=DEV)
    xp = make_plane(x, Ci, 0, True, dt)
    wp = pack_conv_weight(w, Co, Ci, dtype=dt)
    p = torch.empty(B, Co, H, W, device=DEV)
    z = torch.empty(B, Co, H, W, device=DEV)
    K.conv3x3(xp, wp, B, H, W, cin=Ci, bias=bias, sample_bias=sb, act=2, out_f32=p, n_valid=Co, sample_out=z,
              uniforms=u)
    ref = torch.sigmoid(ref_conv(x, w, True) + bias.view(1, -1, 1, 1) + sb.view(B, Co, 1, 1))
    assert report("conv f32 head p", p, ref, 1e-4)
    zr = (u < p).float()
    assert torch.equal(z, zr)
    # eval mode: threshold
    K.conv3x3(xp, wp, B, H, W, cin=Ci, bias=bias, sample_bias=sb, act=2, out_f32=p, n_valid=Co, sample_out=z)
    assert torch.equal(z, (p > 0.5).float())


@pytest.mark.parametrize("dt", FWD_DTYPES)
def test_conv_dgrad_epilogue(dt):
    """dgrad = conv with flipped/transposed weights; epilogue adds a residual plane and gates with lrelu'.
    The gradient planes and the dgrad weight operand are bf16 whatever the forward format (one MMA cannot mix formats);
    the gate is read from a forward plane in `dt`."""
    _setup()
    from scm_gan_b200 import kernels as K
    torch.manual_seed(4)
    B, H, W, Ci, Co = 2, 15, 19, 128, 128
    w = bf(torch.randn(Co, Ci, 3, 3, device=DEV) / (3.0 * Ci ** 0.5))
    dy = bf(torch.randn(B, Co, H, W, device=DEV))
    resid = bf(torch.randn(B, Ci, H, W, device=DEV))
    actv = rnd(torch.randn(B, Ci, H, W, device=DEV), dt)
    for wrap in (True, False):
        dyp = make_plane(dy, Co, 0, wrap)
        rp = make_plane(resid, Ci, 0, wrap)
        ap = make_plane(actv, Ci, 0, wrap, dt)
        wd = pack_conv_weight(w, Ci, Co, dgrad=True)
        out = K.new_plane(B, H, W, Ci, DEV)
        K.conv3x3(dyp, wd, B, H, W, cin=Co, out=out, wrap=wrap, add=rp, gate=ap, dgrad=True)
        # reference: autograd of the forward conv
        xin = torch.zeros(B, Ci, H, W, device=DEV, requires_grad=True)
        y = ref_conv(xin, w, wrap)
        (gx,) = torch.autograd.grad(y, xin, dy)
        ref = (gx + resid) * torch.where(actv > 0, 1.0, 0.01)
        got = plane_interior(out, 0, Ci)
        assert report(f"dgrad wrap={wrap}", got, ref, 4e-3)


@pytest.mark.parametrize("dt", FWD_DTYPES)
@pytest.mark.parametrize("shape", [(3, 64, 64, 128), (4, 15, 19, 64), (2, 16, 16, 128)])
def test_expand_conv_sample_bias_scale_and_gated_dgrad(shape, dt):
    """conv_expand.cuh beyond the plain cases: per-sample bias + per-sample scale (Transition conv1 with the batch
    folded over several spectral-norm calls), and the gated data gradient (16-channel gradient plane -> 64/128
    channels times lrelu'(saved activation), gate tile loaded by TMA)."""
    _setup()
    from scm_gan_b200 import kernels as K
    B, H, W, Co = shape
    Ci = 16
    torch.manual_seed(21)
    x = rnd(torch.randn(B, Ci, H, W, device=DEV), dt)
    w = rnd(torch.randn(Co, Ci, 3, 3, device=DEV) / 12.0, dt)
    sb = torch.randn(B, Co, device=DEV)
    ss = torch.rand(B, device=DEV) + 0.5
    for wrap in (True, False):
        xp = make_plane(x, Ci, 0, wrap, dt)
        wp = pack_conv_weight(w, Co, Ci, dtype=dt)
        out = torch.full((B, H + 2, W + 2, 2 * Co), float("nan"), dtype=dt, device=DEV)
        K.conv3x3(xp, wp, B, H, W, cin=Ci, sample_bias=sb, sample_scale=ss, act=1, out=out, out_c_off=Co, wrap=wrap)
        ref = F.leaky_relu(ref_conv(x, w, wrap) * ss.view(B, 1, 1, 1) + sb.view(B, Co, 1, 1))
        got = plane_interior(out, Co, Co)
        assert report(f"expand fwd {shape} wrap={wrap} {dt}", got, ref, OUT_TOL[dt])
        full = out[:, :, :, Co:].permute(0, 3, 1, 2).float()
        assert torch.equal(full, ref_plane(got, wrap)), "halo mismatch"
        assert torch.isnan(out[:, :, :, :Co].float()).all()
    # gated dgrad: bf16 gradient plane and weights, gate read from a forward plane in `dt`
    dy = bf(torch.randn(B, Ci, H, W, device=DEV))
    wg = bf(torch.randn(Ci, Co, 3, 3, device=DEV) / 12.0)     # Conv2d weight [co=16][ci=Co]
    actv = rnd(torch.randn(B, Co, H, W, device=DEV), dt)
    for wrap in (True, False):
        dyp = make_plane(dy, Ci, 0, wrap)
        ap = make_plane(actv, Co + 16, 16, wrap, dt)
        wd = pack_conv_weight(wg, Co, Ci, dgrad=True)
        out = K.new_plane(B, H, W, Co, DEV)
        K.conv3x3(dyp, wd, B, H, W, cin=Ci, out=out, wrap=wrap, gate=ap, gate_c_off=16, sample_scale=ss, dgrad=True)
        xin = torch.zeros(B, Co, H, W, device=DEV, requires_grad=True)
        (gx,) = torch.autograd.grad(ref_conv(xin, wg, wrap), xin, dy)
        ref = gx * ss.view(B, 1, 1, 1) * torch.where(actv > 0, 1.0, 0.01)
        got = plane_interior(out, 0, Co)
        assert report(f"expand gated dgrad {shape} wrap={wrap} {dt}", got, ref, 4e-3)
        full = out.permute(0, 3, 1, 2).float()
        assert torch.equal(full, ref_plane(got, wrap)), "halo mismatch"


WGRAD_CASES = [
    # (B, H, W, Cin, Cout, wrap)
    (3, 15, 19, 128, 128, True),
    (2, 64, 64, 128, 128, True),
    (2, 15, 19, 256, 128, True),
    (3, 15, 19, 16, 128, False),
    (3, 15, 19, 256, 16, True),
    (2, 16, 16, 128, 16, False),
    (2, 15, 19, 128, 48, False),
    # widths that are multiples of 16 take the row-shifted-reuse kernel (conv_wgrad_v2.cuh)
    (2, 16, 16, 16, 128, False),
    (2, 12, 32, 256, 128, True),
    (3, 10, 48, 64, 128, True),
    # ... and so do other widths, padded to the next multiple of 16 inside the TMA boxes (MiniPacMan: 19 -> 32)
    (2, 9, 37, 128, 128, True),
    (1, 20, 70, 64, 128, False),
    # 16 output channels at the bench plane size: nine taps folded into N (conv_wgrad_narrow.cuh)
    (2, 64, 64, 256, 16, True),
    (1, 33, 70, 128, 16, False),
]


@pytest.mark.parametrize("dt", FWD_DTYPES)
@pytest.mark.parametrize("case", WGRAD_CASES)
def test_wgrad(case, dt):
    """dY is a bf16 gradient plane, X a forward plane in `dt`; with dt = fp16 the kernel rewrites the X tiles to bf16 in
    shared memory (the reference below therefore sees X rounded to fp16 and then to bf16)."""
    _setup()
    from scm_gan_b200 import kernels as K
    B, H, W, Ci, Co, wrap = case
    torch.manual_seed(5)
    x = rnd(torch.randn(B, Ci, H, W, device=DEV), dt)
    dy = bf(torch.randn(B, Co, H, W, device=DEV))
    xp = make_plane(x, Ci, 0, wrap, dt)
    x = bf(x)
    dyp = make_plane(dy, Co, 0, wrap)  # halo deliberately non-zero in wrap mode: must be ignored
    g = torch.zeros(Co, Ci, 3, 3, device=DEV)
    db = torch.zeros(Co, device=DEV)
    K.wgrad(dyp, xp, g, B, H, W, cout=Co, cin=Ci, g_s_co=Ci * 9, g_s_ci=9, db=db)
    w = torch.zeros(Co, Ci, 3, 3, device=DEV, requires_grad=True)
    y = ref_conv(x, w, wrap)
    (gw,) = torch.autograd.grad(y, w, dy)
    assert report(f"wgrad {case}", g, gw, 1e-4)
    # fused bias gradient (all-ones GEMM column in the v2 kernel, column-sum fallback otherwise)
    assert report(f"wgrad bias {case}", db, dy.sum((0, 2, 3)), 1e-4)


def test_spectral_norm_fwd_bwd():
    _setup()
    from scm_gan_b200 import kernels as K
    torch.manual_seed(6)
    shapes = [(128, 21, 3, 3), (128, 128, 3, 3), (128, 256, 3, 3)]
    ws = [torch.randn(s, device=DEV) * 0.05 for s in shapes]
    us = [F.normalize(torch.randn(s[0], device=DEV), dim=0) for s in shapes]
    vs = [F.normalize(torch.randn(s[1] * 9, device=DEV), dim=0) for s in shapes]
    sig = torch.zeros(len(shapes), device=DEV)
    u2 = [u.clone() for u in us]
    v2 = [v.clone() for v in vs]
    usave = [torch.empty_like(u) for u in us]
    vsave = [torch.empty_like(v) for v in vs]
    K.spectral_norm_fwd([(w, u2[i], v2[i], sig[i:i + 1], usave[i], vsave[i]) for i, w in enumerate(ws)])
    for i, w in enumerate(ws):
        wm = w.view(w.shape[0], -1)
        v = wm.t().mv(us[i]); v = v / (v.norm() + 1e-12)
        u = wm.mv(v); u = u / (u.norm() + 1e-12)
        s = u.dot(wm.mv(v))
        assert report(f"sn v {i}", v2[i], v, 1e-5)
        assert report(f"sn u {i}", u2[i], u, 1e-5)
        assert abs(sig[i].item() - s.item()) <= 1e-5 * abs(s.item())
        assert torch.equal(usave[i], u2[i]) and torch.equal(vsave[i], v2[i])
        # backward
        g = torch.randn_like(w)
        wb = w.clone().requires_grad_(True)
        sigma = u.dot(wb.view(wb.shape[0], -1).mv(v))
        (ref,) = torch.autograd.grad(wb / sigma.expand_as(wb), wb, g)
        dot = torch.zeros(1, device=DEV)
        out = torch.empty_like(w)
        K.spectral_norm_bwd([(g, w, usave[i], vsave[i], sig[i:i + 1], dot, out)])
        assert report(f"sn bwd {i}", out, ref, 1e-4)


def test_colsum_action_bce_adam():
    _setup()
    from scm_gan_b200 import kernels as K
    torch.manual_seed(7)
    B, H, W, Cc = 3, 15, 19, 128
    x = bf(torch.randn(B, Cc, H, W, device=DEV))
    xp = make_plane(x, Cc + 16, 16, True)
    S = torch.zeros(B, Cc, device=DEV)
    db = torch.zeros(Cc, device=DEV)
    K.plane_colsum(xp, 16, Cc, B, H, W, S=S, db=db)
    assert report("colsum S", S, x.sum((2, 3)), 1e-5)
    assert report("colsum db", db, x.sum((0, 2, 3)), 1e-5)

    # folded action channels
    Lz, A = 16, 5
    wbar = torch.randn(128, Lz + A, 3, 3, device=DEV)
    sigma = torch.tensor([1.7], device=DEV)
    bias = torch.randn(128, device=DEV)
    act = torch.eye(A, device=DEV)[torch.randint(A, (B,), device=DEV)]
    sb = torch.empty(B, 128, device=DEV)
    K.action_bias(wbar, sigma, bias, act, Lz, sb)
    ref = bias + act @ (wbar[:, Lz:].sum((2, 3)) / sigma).t()
    assert report("action_bias", sb, ref, 1e-5)
    g = torch.zeros_like(wbar)
    K.action_wgrad(S, act, Lz, g)
    refg = (S.t() @ act).view(128, A, 1, 1).expand(128, A, 3, 3)
    assert report("action_wgrad", g[:, Lz:], refg, 1e-5)
    assert g[:, :Lz].abs().max().item() == 0

    # fused sigmoid + BCE + masked mean, forward value and gradient
    logits = torch.randn(B, 3, H, W, device=DEV) * 3
    states = (torch.rand(B, 4, 3, H, W, device=DEV) < 0.15).float()
    tgt = states[:, 2]
    mask = torch.tensor([1.0, 0.0, 1.0], device=DEV)
    loss = torch.zeros(1, device=DEV)
    dx = torch.empty_like(logits)
    K.bce_logits(logits, tgt, mask, loss, dx)
    lg = logits.clone().requires_grad_(True)
    rl = (F.binary_cross_entropy(torch.sigmoid(lg), tgt, reduction="none").mean(-1).mean(-1).mean(-1) * mask).mean()
    rl.backward()
    assert abs(loss.item() - rl.item()) <= 1e-5 * abs(rl.item())
    assert report("bce dx", dx, lg.grad, 1e-5)
    # sequence form: the decodes of T rollout steps as one batch (t-major logits, [B, T] windows of states / masks)
    import scm_gan_b200.ops  # noqa: F401
    Tn = 3
    lt = (torch.randn(Tn * B, 3, H, W, device=DEV) * 3).requires_grad_(True)
    mbt = (torch.rand(B, 4, device=DEV) > 0.3).float()
    refs = [(F.binary_cross_entropy(torch.sigmoid(lt[t * B:(t + 1) * B]), states[:, 1 + t], reduction="none")
             .mean(-1).mean(-1).mean(-1) * mbt[:, t]).mean() for t in range(Tn)]
    wts = torch.tensor([1.0, 2.0, 0.5], device=DEV)
    (gr,) = torch.autograd.grad((torch.stack(refs) * wts).sum(), lt)
    terms = torch.ops.scmgan.bce_logits_seq(lt, states[:, 1:1 + Tn], mbt[:, :Tn])[0]
    (gg,) = torch.autograd.grad((terms * wts).sum(), lt)
    assert report("bce seq terms", terms, torch.stack(refs).detach(), 1e-5)
    assert report("bce seq grad", gg, gr, 1e-5)

    # fused clip + Adam vs torch.optim.Adam after clip_grad_value_
    ps = [torch.randn(1000, device=DEV), torch.randn(77, 3, device=DEV)]
    gs = [torch.randn_like(p) * 0.3 for p in ps]
    ref_ps = [p.clone().requires_grad_(True) for p in ps]
    opt = torch.optim.Adam(ref_ps, lr=1e-4)
    ms = [torch.zeros_like(p) for p in ps]
    vv = [torch.zeros_like(p) for p in ps]
    for step in range(1, 4):
        for rp, g in zip(ref_ps, gs):
            rp.grad = g.clone()
        torch.nn.utils.clip_grad_value_(ref_ps, 0.1)
        opt.step()
        K.clip_adam([(p, g, m, v, 0.1) for p, g, m, v in zip(ps, gs, ms, vv)], 1e-4, 0.9, 0.999, 1e-8, step)
    for p, rp in zip(ps, ref_ps):
        assert report("adam", p, rp.detach(), 1e-6)


if __name__ == "__main__":
    import sys
    import traceback
    fails = 0

    def run(fn, *a):
        global fails
        try:
            fn(*a)
            torch.cuda.synchronize()
        except Exception:
            fails += 1
            traceback.print_exc()
            print(f"FAILED {fn.__name__} {a}", flush=True)
            try:
                torch.cuda.synchronize()
            except Exception as e:  # sticky CUDA error: stop
                print("sticky CUDA error, aborting:", e)
                sys.exit(2)

    only = sys.argv[1] if len(sys.argv) > 1 else ""
    if only in ("", "pack"):
        for dt in FWD_DTYPES:
            run(test_pack_nchw, False, dt)
            run(test_pack_nchw, True, dt)
    if only in ("", "conv"):
        for dt in FWD_DTYPES:
            for c in CONV_CASES:
                run(test_conv_fwd_plane, c, dt)
            run(test_conv_sample_bias_and_f32_head, dt)
            run(test_conv_dgrad_epilogue, dt)
    if only in ("", "wgrad"):
        for dt in FWD_DTYPES:
            for c in WGRAD_CASES:
                run(test_wgrad, c, dt)
    if only in ("", "misc"):
        run(test_spectral_norm_fwd_bwd)
        run(test_colsum_action_bce_adam)
    print("failures:", fails)
    sys.exit(1 if fails else 0)


def test_cf_losses_and_transition_tail():
    """Fused counterfactual-loss kernels (reference main.py:258-262, 279-283) and the stand-alone Transition tail."""
    _setup()
    import scm_gan_b200.ops  # noqa: F401
    from scm_gan_b200 import kernels as K
    torch.manual_seed(8)
    B, Lz, H, W = 5, 16, 15, 19
    za = (torch.rand(B, Lz, H, W, device=DEV) < 0.5).float().requires_grad_(True)
    zb = torch.rand(B, Lz, H, W, device=DEV).requires_grad_(True)
    mask = torch.tensor([1.0, 1.0, 0.0, 1.0, 1.0], device=DEV)
    unsw = (torch.rand(B, Lz, device=DEV) < 0.8).float()
    lam = 0.01
    for mode in (0, 1):
        if mode == 0:
            ref = torch.abs(za - zb).mean(-1).mean(-1) * unsw
            ref = lam * torch.mean(ref.mean(-1) * mask)
        else:
            ref = -torch.log(torch.abs(za - zb).mean(-1).mean(-1).mean(-1) + 0.001)
            ref = lam * torch.mean(ref * mask)
        ga, gb = torch.autograd.grad(ref * 3.0, (za, zb))
        got = torch.ops.scmgan.cf_loss(za, zb, unsw if mode == 0 else None, mask, mode, lam)[0]
        ha, hb = torch.autograd.grad(got * 3.0, (za, zb))
        assert abs(got.item() - ref.item()) <= 1e-5 * abs(ref.item()) + 1e-9, mode
        assert report(f"cf_loss mode {mode} dza", ha, ga, 1e-5)
        assert report(f"cf_loss mode {mode} dzb", hb, gb, 1e-5)
    x = torch.randn(B, Lz, H, W, device=DEV)
    u = torch.rand_like(x)
    p, z = torch.empty_like(x), torch.empty_like(x)
    K.transition_tail(x, u, p, z)
    assert report("tail p", p, torch.sigmoid(x), 1e-6)
    assert torch.equal(z, (u < p).float())
    K.transition_tail(x, None, p, z)
    assert torch.equal(z, (p > 0.5).float())


def test_in_kernel_philox_bernoulli():
    """Bernoulli head with the in-kernel Philox4x32-10 stream: calibrated, reproducible from {seed, offset}, offset
    advanced on the device (graph-replay safe)."""
    _setup()
    from scm_gan_b200 import kernels as K
    torch.manual_seed(9)
    B, H, W, Ci, Co = 8, 32, 32, 64, 16
    x = bf(torch.randn(B, Ci, H, W, device=DEV))
    w = bf(torch.randn(Co, Ci, 3, 3, device=DEV) / (1.5 * Ci ** 0.5))
    xp = make_plane(x, Ci, 0, True)
    wp = pack_conv_weight(w, Co, Ci)
    n = B * Co * H * W

    def draw(state):
        p = torch.empty(B, Co, H, W, device=DEV)
        z = torch.empty(B, Co, H, W, device=DEV)
        K.conv3x3(xp, wp, B, H, W, cin=Ci, act=2, out_f32=p, n_valid=Co, sample_out=z, rng_state=state)
        return p, z

    s1 = torch.tensor([1234, 0], dtype=torch.int64, device=DEV)
    p, z1 = draw(s1)
    assert s1.tolist() == [1234, (n + 3) // 4]
    assert set(z1.unique().tolist()) <= {0.0, 1.0}
    # calibration: E[z] = p.  Overall and per probability bin (4 sigma)
    assert abs(z1.mean().item() - p.mean().item()) < 4 * 0.5 / n ** 0.5
    for lo in (0.0, 0.2, 0.4, 0.6, 0.8):
        sel = (p >= lo) & (p < lo + 0.2)
        m = int(sel.sum())
        if m > 1000:
            assert abs(z1[sel].mean().item() - p[sel].mean().item()) < 4 * 0.5 / m ** 0.5, lo
    _, z2 = draw(s1)                      # continues the stream: independent draw
    agree = (z1 == z2).float().mean().item()
    expect = (p * p + (1 - p) * (1 - p)).mean().item()
    assert abs(agree - expect) < 0.01
    s3 = torch.tensor([1234, 0], dtype=torch.int64, device=DEV)
    _, z3 = draw(s3)                      # same {seed, offset}: same sample
    assert torch.equal(z1, z3)
    _, z4 = draw(torch.tensor([99, 0], dtype=torch.int64, device=DEV))
    assert not torch.equal(z1, z4)


def test_philox_uniform_stream_matches_in_kernel_head():
    """scmgan_philox_uniform writes the very uniforms the in-kernel Bernoulli head draws: sampling with the filled
    tensor passed as `uniforms` is bit-identical to sampling with `rng_state`, and the offsets advance alike."""
    _setup()
    from scm_gan_b200 import kernels as K
    torch.manual_seed(10)
    B, H, W, Ci, Co = 4, 16, 16, 64, 16
    x = bf(torch.randn(B, Ci, H, W, device=DEV))
    w = bf(torch.randn(Co, Ci, 3, 3, device=DEV) / (1.5 * Ci ** 0.5))
    xp = make_plane(x, Ci, 0, True)
    wp = pack_conv_weight(w, Co, Ci)
    s_a = torch.tensor([77, 5], dtype=torch.int64, device=DEV)
    s_b = s_a.clone()
    p1, z1 = torch.empty(B, Co, H, W, device=DEV), torch.empty(B, Co, H, W, device=DEV)
    K.conv3x3(xp, wp, B, H, W, cin=Ci, act=2, out_f32=p1, n_valid=Co, sample_out=z1, rng_state=s_a)
    u = torch.empty(B, Co, H, W, device=DEV)
    K.philox_uniform(u, s_b)
    assert s_a.tolist() == s_b.tolist()
    assert 0.0 < u.min().item() and u.max().item() < 1.0
    assert abs(u.mean().item() - 0.5) < 4 * (1 / 12) ** 0.5 / u.numel() ** 0.5
    p2, z2 = torch.empty_like(p1), torch.empty_like(z1)
    K.conv3x3(xp, wp, B, H, W, cin=Ci, act=2, out_f32=p2, n_valid=Co, sample_out=z2, uniforms=u)
    assert torch.equal(p1, p2) and torch.equal(z1, z2)
    # odd length: the tail of the last Philox block is dropped, the offset still advances by whole blocks
    v = torch.empty(10, device=DEV)
    s_c = torch.tensor([77, 5], dtype=torch.int64, device=DEV)
    K.philox_uniform(v, s_c)
    assert torch.equal(v, u.flatten()[:10]) and s_c.tolist() == [77, 8]


def test_masked_mse():
    """Reward regression term (reference main.py:182-186): value and gradient of the fused kernel vs torch."""
    _setup()
    torch.manual_seed(12)
    B, Hn, R = 32, 10, 3
    rewards = torch.randn(B, Hn, R, device=DEV)
    masks = (torch.rand(B, Hn - 1, device=DEV) > 0.3).float()
    pred = torch.randn(B, R, device=DEV, requires_grad=True)
    ref = 0.37 * torch.mean(torch.mean((pred - rewards[:, 4]) ** 2, dim=1) * masks[:, 3])
    (gref,) = torch.autograd.grad(ref, pred)
    p2 = pred.detach().clone().requires_grad_(True)
    got = torch.ops.scmgan.masked_mse(p2, rewards[:, 4], masks[:, 3], 0.37, None)[0]
    (ggot,) = torch.autograd.grad(got * 2.0, p2)
    assert report("masked mse", got, ref, 1e-6)
    assert report("masked mse grad", ggot, 2.0 * gref, 1e-6)
    # theta as a device scalar (one CUDA graph for every training iteration) + the unscaled value the reference logs
    theta = torch.tensor(0.25, device=DEV)
    p3 = pred.detach().clone().requires_grad_(True)
    out = torch.ops.scmgan.masked_mse(p3, rewards[:, 4], masks[:, 3], 0.37, theta)
    (g3,) = torch.autograd.grad(out[0], p3)
    assert report("masked mse (device theta)", out[0], 0.25 * ref, 1e-6)
    assert report("masked mse grad (device theta)", g3, 0.25 * gref, 1e-6)
    assert report("masked mse raw", out[2], ref / 0.37, 1e-6)
    # sequence form: T steps in one launch, targets / masks addressed as [B, T] windows of the batch tensors
    T = 5
    pt = torch.randn(T * B, R, device=DEV, requires_grad=True)
    refs = [torch.mean(torch.mean((pt[t * B:(t + 1) * B] - rewards[:, 2 + t]) ** 2, dim=1) * masks[:, 1 + t])
            for t in range(T)]
    ref_tot = 0.37 * 0.25 * sum(refs)
    (gref,) = torch.autograd.grad(ref_tot, pt)
    out = torch.ops.scmgan.masked_mse_seq(pt, rewards[:, 2:2 + T], masks[:, 1:1 + T], 0.37, theta)
    (gg,) = torch.autograd.grad(out[0], pt)
    assert report("masked mse seq", out[0], ref_tot, 1e-6)
    assert report("masked mse seq raw", out[2], torch.stack(refs), 1e-6)
    assert report("masked mse seq grad", gg, gref, 1e-6)


@pytest.mark.parametrize("shape", [(2, 8, 6, 7), (3, 16, 12, 9), (2, 32, 16, 16)])
def test_csrn_native_sweeps_vs_oracle(shape):
    """scmgan_gru_conv_sweep_fwd/bwd (interface-only layer CSRN, reference spatial_recurrent.py:61-114): output and
    every gradient of the drop-in module on CUDA (hand-written sweep kernels) against the fp32 oracle restatement
    differentiated by torch autograd."""
    _setup()
    import sys
    from oracle import restated as R
    from scm_gan_b200.train_step import import_dropin_models
    import_dropin_models()
    CSRN = sys.modules["spatial_recurrent"].CSRN
    B, C, H, W = shape
    torch.manual_seed(3)
    net = CSRN(C).to(DEV)
    with torch.no_grad():  # the reference's N(0, channels) init saturates every gate; use a scale with live gradients
        for n_, p in net.named_parameters():
            if "combine" not in n_:
                p.normal_(0, 0.5 / C ** 0.5)
    x = (torch.randn(B, C, H, W, device=DEV) * 0.7).requires_grad_(True)
    y = net(x)
    gy = torch.randn_like(y)
    y.backward(gy)
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in net.state_dict().items()}
    x2 = x.detach().clone().requires_grad_(True)
    yr = R.csrn_forward(sd, x2)
    yr.backward(gy)
    assert report(f"csrn fwd {shape}", y, yr, 1e-5)
    assert report(f"csrn dx {shape}", x.grad, x2.grad, 1e-4)
    for k, p in net.named_parameters():
        ref = sd[k].grad
        if ref is None or ref.abs().max().item() == 0:   # rnn_left / conv_left: overwritten by the reference quirk
            assert p.grad is None or p.grad.abs().max().item() == 0, k
            continue
        assert report(f"csrn d{k} {shape}", p.grad, ref, 1e-4)


@pytest.mark.parametrize("shape", [(2, 6, 12, 12, 16), (3, 14, 16, 16, 32), (2, 30, 10, 13, 24)])
def test_coordconv_native_vs_torch(shape):
    """CoordConv2d 3x3/s1/p1 on the hand-written path against the same layer evaluated by torch in fp32 (coordinates
    concatenated as in reference coordconv.py:10-15); 16-bit operands: 1e-2 relative.  The first two shapes (<= 14 data
    channels) take the in-tile path - coordinates generated by the conv's producer while it stages the im2col tile
    (scmgan_conv_desc::coord_c1), their weight gradient from scmgan_coord_wgrad; the third (30 data channels, TMA-fed
    tiles) materialises them with scmgan_pack_coords."""
    _setup()
    import sys
    from scm_gan_b200.train_step import import_dropin_models
    import_dropin_models()
    CoordConv2d = sys.modules["coordconv"].CoordConv2d
    B, C, H, W, Co = shape
    torch.manual_seed(4)
    net = CoordConv2d(C + 2, Co, 3, padding=1).to(DEV)
    x = torch.randn(B, C, H, W, device=DEV, requires_grad=True)
    assert net._native_ok(x)
    from scm_gan_b200 import engine
    assert engine.coords_in_tile((C + 2 + 15) // 16 * 16, W) == (C + 2 <= 16)
    y = net(x)
    gy = torch.randn_like(y)
    y.backward(gy)
    x2 = x.detach().clone().requires_grad_(True)
    cx, cy = net.coordinates(H, W, DEV, torch.float32)
    xin = torch.cat([x2, cx.expand(B, -1, -1, -1), cy.expand(B, -1, -1, -1)], dim=1)
    w2 = net.conv.weight.detach().clone().requires_grad_(True)
    b2 = net.conv.bias.detach().clone().requires_grad_(True)
    yr = torch.nn.functional.conv2d(xin, w2, b2, padding=1)
    yr.backward(gy)
    assert report(f"coordconv fwd {shape}", y, yr, 1e-2)
    assert report(f"coordconv dx {shape}", x.grad, x2.grad, 1e-2)
    assert report(f"coordconv dW {shape}", net.conv.weight.grad, w2.grad, 1e-2)
    assert report(f"coordconv db {shape}", net.conv.bias.grad, b2.grad, 1e-2)
    if shape == (2, 6, 12, 12, 16):  # the layer recorded from the unmodified reference (tests/golden/layers.pt)
        g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "layers.pt"), weights_only=False)["coordconv"]
        ref = CoordConv2d(6 + 2, 16, 3, padding=1).to(DEV)
        ref.load_state_dict(g["state"])
        with torch.no_grad():
            assert report("coordconv vs reference golden", ref(g["x"].to(DEV)), g["y"].to(DEV), 1e-2)
