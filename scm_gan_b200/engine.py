"""Kernel sequencing for the three convolutional networks of the world model (host side, no arithmetic).

Each function enqueues the hand-written kernels (scm_gan_b200.kernels -> C ABI) for one module forward or
backward on torch's current stream and returns the tensors autograd has to keep.  Activations stay in bf16
"planes" ([B, H+2, W+2, C], see include/scmgan.h) between layers; fp32 NCHW exists only at the module boundary.

Reference semantics mirrored here (file:line in LilJing/scm-gan):
  Encoder.forward      models.py:139-157     zero padding, SN convs 1-3, plain conv4, sigmoid
  Transition.forward   models.py:59-119      legacy circular pad-1, SN convs 1-5, plain conv6, sigmoid, Bernoulli
  Decoder.forward      models.py:270-291     two ConvTranspose2d(3x3, s1, p1), latent-group sum
  SpectralNorm         spectral_normalization.py:23-35  (power iteration per call; backward uses the u, v held by
                       the module at backward time, see DESIGN.md "SpectralNorm backward")
"""
import torch

from . import kernels as K
from .kernels import ACT_LRELU, ACT_NONE, ACT_SIGMOID

HID = 128  # hidden width of Encoder / Transition (reference models.py:51-55, 129-133)


def _r16(c):
    return (c + 15) // 16 * 16


def _conv2d_fwd_job(w, out, sigma=None, k_valid=None, n_pad=None, k_pad=None):
    """nn.Conv2d weight [Co, Ci, 3, 3] -> forward GEMM operand [9][n_pad][k_pad] (n = co, k = ci)."""
    co, ci = w.shape[0], w.shape[1]
    return dict(w=w, out=out, sigma=sigma, n_pad=out.shape[1], k_pad=out.shape[2], n_valid=co,
                k_valid=ci if k_valid is None else k_valid, s_n=ci * 9, s_k=9, flip=0)


def _conv2d_dgrad_job(w, out, sigma=None, ci_begin=0, ci_count=None, co_valid=None):
    """nn.Conv2d weight -> dgrad operand [9][n_pad][k_pad] with n = ci (window), k = co, taps flipped."""
    co, ci = w.shape[0], w.shape[1]
    ci_count = ci - ci_begin if ci_count is None else ci_count
    # window over ci: shift the base pointer by ci_begin * 9 elements via a view
    wv = w.view(co, ci * 9)[:, ci_begin * 9:]
    return dict(w=wv, out=out, sigma=sigma, n_pad=out.shape[1], k_pad=out.shape[2], n_valid=ci_count,
                k_valid=co if co_valid is None else co_valid, s_n=9, s_k=ci * 9, flip=1)


def _convT_fwd_job(w, out):
    """nn.ConvTranspose2d weight [Ci, Co, 3, 3] (stride 1, pad 1) -> equivalent correlation operand (flipped)."""
    ci, co = w.shape[0], w.shape[1]
    return dict(w=w, out=out, sigma=None, n_pad=out.shape[1], k_pad=out.shape[2], n_valid=co, k_valid=ci, s_n=9,
                s_k=co * 9, flip=1)


def _convT_dgrad_job(w, out):
    ci, co = w.shape[0], w.shape[1]
    return dict(w=w, out=out, sigma=None, n_pad=out.shape[1], k_pad=out.shape[2], n_valid=ci, k_valid=co,
                s_n=co * 9, s_k=9, flip=0)


# ------------------------------------------------------------------------------------------------------------
# Transition
# ------------------------------------------------------------------------------------------------------------
def spectral_norm_update(wbar, u, v):
    """One power iteration for every wrapped conv of a module (u, v updated in place); returns sigma [n]."""
    sigma = torch.empty(len(wbar), dtype=torch.float32, device=wbar[0].device)
    K.spectral_norm_fwd([(wbar[i], u[i], v[i], sigma[i:i + 1], None, None) for i in range(len(wbar))])
    return sigma


def _segment_ratio(sigma, B):
    """sigma [S, n_layers] -> rho [n_layers, B] with rho[l, b] = sigma[0, l] / sigma[segment(b), l], or None for S = 1.
    The batch is S equal segments whose spectral-norm calls differ (main rollout step + counterfactual rollouts folded
    into one batch); the weights are packed for segment 0's sigma and rho rescales the other segments' rows."""
    S = sigma.shape[0]
    if S == 1:
        return None
    assert B % S == 0
    return (sigma[0:1] / sigma).t().repeat_interleave(B // S, dim=1).contiguous()


def transition_forward(z, a, wbar, bias, sigma, w6, b6, uniforms, training, rng_state=None):
    """z [B,L,H,W] fp32, a [B,A] fp32.  wbar/bias: lists for conv1..conv5; sigma [5] from spectral_norm_update, or
    [S, 5] when the batch consists of S equal segments that belong to S different calls of the reference (see
    _segment_ratio).  Returns (z_next, p, saved) with saved = [zin, buf6, buf5, act3, wd..., rho] for backward."""
    dev = z.device
    B, L, H, W = z.shape
    Lp = _r16(L)
    A = a.shape[1]
    sigma = sigma.view(-1, 5)
    S = sigma.shape[0]
    sig0 = sigma[0]
    rho = _segment_ratio(sigma, B)
    rs = (lambda l: None) if rho is None else (lambda l: rho[l])

    # packed operands: forward [9][Cout][Cin] and dgrad [9][Cin][Cout] (1/sigma folded in)
    wf = [K.packed_weight(HID, Lp, dev), K.packed_weight(HID, HID, dev), K.packed_weight(HID, HID, dev),
          K.packed_weight(HID, HID, dev), K.packed_weight(HID, 2 * HID, dev), K.packed_weight(Lp, 2 * HID, dev)]
    # dgrad operands.  The two gradients that meet at a skip connection are one GEMM over concatenated K:
    #   d act2 = [d pre3 | d pre5] x [Wd3 ; Wd5(skip half)]     (K = 128 + 128)
    #   d act1 = [d pre2 | d pre6] x [Wd2 ; Wd6(skip half)]     (K = 128 + 64, the 16 latent gradients zero-padded)
    LS = 64  # channel slot of d pre6 inside the [d pre2 | d pre6] plane
    assert Lp <= LS
    G16 = K.GRAD_DTYPE  # dgrad multiplies bf16 gradient planes: its weight operands are packed in the same format
    wd = [K.packed_weight(Lp, HID, dev, G16), K.packed_weight(HID, HID + LS, dev, G16),
          K.packed_weight(HID, 2 * HID, dev, G16), K.packed_weight(HID, HID, dev, G16),
          K.packed_weight(HID, HID, dev, G16), K.packed_weight(HID, Lp, dev, G16)]
    jobs = [_conv2d_fwd_job(wbar[0], wf[0], sig0[0:1], k_valid=L)]
    jobs += [_conv2d_fwd_job(wbar[i], wf[i], sig0[i:i + 1]) for i in range(1, 5)]
    jobs += [_conv2d_fwd_job(w6, wf[5])]
    jobs += [_conv2d_dgrad_job(wbar[0], wd[0], sig0[0:1], 0, L)]
    jobs += [_conv2d_dgrad_job(wbar[1], wd[1][:, :, :HID], sig0[1:2]),                       # conv2
             _conv2d_dgrad_job(w6, wd[1][:, :, HID:], None, HID, HID, co_valid=L)]              # conv6, skip half
    jobs += [_conv2d_dgrad_job(wbar[2], wd[2][:, :, :HID], sig0[2:3]),                       # conv3
             _conv2d_dgrad_job(wbar[4], wd[2][:, :, HID:], sig0[4:5], HID, HID)]               # conv5, skip half
    jobs += [_conv2d_dgrad_job(wbar[3], wd[3], sig0[3:4])]                                   # conv4
    jobs += [_conv2d_dgrad_job(wbar[4], wd[4], sig0[4:5], 0, HID)]                           # conv5, act4 half
    jobs += [_conv2d_dgrad_job(w6, wd[5], None, 0, HID, co_valid=L)]                          # conv6, act5 half
    K.pack_weights(jobs)

    # folded action channels: per-sample bias with the sigma of the sample's own segment
    sbias = torch.empty((B, HID), dtype=torch.float32, device=dev)
    Bs = B // S
    for s in range(S):
        K.action_bias(wbar[0], sigma[s, 0:1], bias[0], a[s * Bs:(s + 1) * Bs], L, sbias[s * Bs:(s + 1) * Bs])

    zin = K.fwd_plane(B, H, W, Lp, dev)
    K.pack_nchw(z, zin, wrap=True)
    buf6 = K.fwd_plane(B, H, W, 2 * HID, dev)  # [act5 | act1]  = input of conv6 (models.py:101)
    buf5 = K.fwd_plane(B, H, W, 2 * HID, dev)  # [act4 | act2]  = input of conv5 (models.py:95)
    act3 = K.fwd_plane(B, H, W, HID, dev)
    cv = dict(act=ACT_LRELU, wrap=True)
    K.conv3x3(zin, wf[0], B, H, W, cin=Lp, sample_bias=sbias, sample_scale=rs(0), out=buf6, out_c_off=HID, **cv)  # conv1 -> skip1
    K.conv3x3(buf6, wf[1], B, H, W, cin=HID, x_c_off=HID, bias=bias[1], sample_scale=rs(1), out=buf5, out_c_off=HID, **cv)  # conv2 -> skip2
    K.conv3x3(buf5, wf[2], B, H, W, cin=HID, x_c_off=HID, bias=bias[2], sample_scale=rs(2), out=act3, **cv)  # conv3
    K.conv3x3(act3, wf[3], B, H, W, cin=HID, bias=bias[3], sample_scale=rs(3), out=buf5, out_c_off=0, **cv)  # conv4
    K.conv3x3(buf5, wf[4], B, H, W, cin=2 * HID, bias=bias[4], sample_scale=rs(4), out=buf6, out_c_off=0, **cv)  # conv5
    p = torch.empty((B, L, H, W), dtype=torch.float32, device=dev)
    zn = torch.empty((B, L, H, W), dtype=torch.float32, device=dev)
    if training and uniforms is None and rng_state is not None:
        # the module's Philox stream, generated in its own pass (same numbers as the in-kernel head would draw,
        # ~10x cheaper than evaluating Philox inside the conv epilogue)
        uniforms = torch.empty((B, L, H, W), dtype=torch.float32, device=dev)
        K.philox_uniform(uniforms, rng_state)
    K.conv3x3(buf6, wf[5], B, H, W, cin=2 * HID, bias=b6, act=ACT_SIGMOID, out_f32=p, n_valid=L, sample_out=zn,
              uniforms=uniforms if training else None)                                              # conv6 + head
    return zn, p, [zin, buf6, buf5, act3] + wd + ([rho] if rho is not None else [])


def transition_backward(dz_next, p, a, saved, wbar, sigma, u, v, w6, sink=None):
    """Backward of transition_forward.  dz_next: gradient w.r.t. the sampled state (straight-through => w.r.t. p,
    reference models.py:38-40).  Returns (dz, [dWbar1..5], [db1..5], dW6, db6).
    sink = [gWbar1..5, gb1..5, gW6, gb6] (entries may be None): the kernels ADD that parameter's gradient into the
    given buffer (the parameter's .grad) and None is returned in its place; the weights are shared by every unrolled
    step, so this replaces one AccumulateGrad add per parameter and step.

    Segments (sigma [S, 5], S > 1): every gradient plane of a spectral-normalised layer l holds
    rho_l[b] * d pre_l (the dgrad epilogue that writes it applies rho), so that the dgrad through layer l with the
    weights packed for sigma[0] yields exactly d pre_l * (Wbar/sigma_s)^T for a sample of segment s.  The weight
    gradients are taken per segment on those planes, G'_s = rho_s G_s, and
        dWbar += G'_s/sigma_0 - (<G'_s, Wbar>/(sigma_0 sigma_s)) u v^T   ( = G_s/sigma_s - (<G_s,Wbar>/sigma_s^2) u v^T )
    while bias gradients are rescaled by 1/rho_s."""
    zin, buf6, buf5, act3 = saved[:4]
    wd = saved[4:10]
    rho = saved[10] if len(saved) > 10 else None
    rs = (lambda l: None) if rho is None else (lambda l: rho[l])
    dev = dz_next.device
    B, L, H, W = dz_next.shape
    Lp = zin.shape[3]
    A = a.shape[1]
    sigma = sigma.view(-1, 5)
    S = sigma.shape[0]
    Bs = B // S
    # gradients w.r.t. the *normalised* weights / biases, one flat zeroed buffer
    shapes = [tuple(w.shape) for w in wbar]
    sizes = [w.numel() for w in wbar]
    nb = 5 * HID
    per_seg = sum(sizes) + nb + Bs * HID + 8
    flat = torch.zeros(S * per_seg + w6.numel() + Lp, dtype=torch.float32, device=dev)
    Gs, dbs, S1s, dots = [], [], [], []
    for s in range(S):
        off = s * per_seg
        G = []
        for sh, n in zip(shapes, sizes):
            G.append(flat[off:off + n].view(sh))
            off += n
        Gs.append(G)
        dbs.append([flat[off + i * HID: off + (i + 1) * HID] for i in range(5)])
        off += nb
        S1s.append(flat[off: off + Bs * HID].view(Bs, HID))
        off += Bs * HID
        dots.append(flat[off: off + 8])
    G6 = flat[S * per_seg: S * per_seg + w6.numel()].view(w6.shape)
    db6 = flat[S * per_seg + w6.numel():]
    sink = list(sink) if sink is not None else [None] * 12
    gw, gb, gw6, gb6 = sink[0:5], sink[5:10], sink[10], sink[11]
    if gw6 is not None:
        G6 = gw6
    if S == 1:  # the bias sums go straight into the caller's buffers
        dbs[0] = [dbs[0][i] if gb[i] is None else gb[i] for i in range(5)]
    if gb6 is not None and Lp == L:
        db6 = gb6

    def seg(plane, s):
        return plane if S == 1 else plane[s * Bs:(s + 1) * Bs]

    # Gradient planes.  DB = [d pre2 | d pre6 (Lp channels, zero-padded to LS)], DA = [d pre3 | d pre5]: each pair is
    # the K-concatenated input of one dgrad GEMM (see transition_forward), so the partial sums of the skip
    # connections never go through memory.
    LS = wd[1].shape[2] - HID
    dr = K.DeferredReduces(dev)  # split-K reductions run on a side stream behind the next conv
    DB = K.new_plane(B, H, W, HID + LS, dev)
    DA = K.new_plane(B, H, W, 2 * HID, dev)
    # d pre-activation of conv6: dz * p * (1 - p)
    K.pack_nchw(dz_next, DB, c_off=HID, c_pad=LS, wrap=True, sig=p)
    cin6 = 2 * HID
    K.wgrad(DB, buf6, G6, B, H, W, cout=Lp, cin=cin6, dy_c_off=HID, g_s_co=cin6 * 9, g_s_ci=9, co_valid=L, defer=dr)
    with dr.side_section():
        K.plane_colsum(DB, HID, Lp, B, H, W, db=db6)
    # (the dgrad operands wd were packed by this call's forward: stable long before the backward runs, which autograd
    # drives from its own thread, where the library's launch bookkeeping cannot see the pack)
    dg = dict(wrap=True, dgrad=True, weights_stable=True)
    K.conv3x3(DB, wd[5], B, H, W, cin=Lp, x_c_off=HID, out=DA, out_c_off=HID, gate=buf6, gate_c_off=0,
              sample_scale=rs(4), **dg)                                                                     # d pre5
    for s in range(S):
        K.wgrad(seg(DA, s), seg(buf5, s), Gs[s][4], Bs, H, W, cout=HID, cin=cin6, dy_c_off=HID, g_s_co=cin6 * 9,
                g_s_ci=9, db=dbs[s][4], defer=dr)
    d4 = K.new_plane(B, H, W, HID, dev)
    K.conv3x3(DA, wd[4], B, H, W, cin=HID, x_c_off=HID, out=d4, gate=buf5, gate_c_off=0, sample_scale=rs(3), **dg)  # d pre4
    for s in range(S):
        K.wgrad(seg(d4, s), seg(act3, s), Gs[s][3], Bs, H, W, cout=HID, cin=HID, g_s_co=HID * 9, g_s_ci=9,
                db=dbs[s][3], defer=dr)
    K.conv3x3(d4, wd[3], B, H, W, cin=HID, out=DA, out_c_off=0, gate=act3, sample_scale=rs(2), **dg)        # d pre3
    for s in range(S):
        K.wgrad(seg(DA, s), seg(buf5, s), Gs[s][2], Bs, H, W, cout=HID, cin=HID, x_c_off=HID, g_s_co=HID * 9,
                g_s_ci=9, db=dbs[s][2], defer=dr)
    K.conv3x3(DA, wd[2], B, H, W, cin=2 * HID, out=DB, out_c_off=0, gate=buf5, gate_c_off=HID, sample_scale=rs(1),
              **dg)                                                                                         # d pre2
    for s in range(S):
        K.wgrad(seg(DB, s), seg(buf6, s), Gs[s][1], Bs, H, W, cout=HID, cin=HID, x_c_off=HID, g_s_co=HID * 9,
                g_s_ci=9, db=dbs[s][1], defer=dr)
    d1 = K.new_plane(B, H, W, HID, dev)
    K.conv3x3(DB, wd[1], B, H, W, cin=HID + LS, out=d1, gate=buf6, gate_c_off=HID, sample_scale=rs(0), **dg)  # d pre1
    c1 = L + A
    for s in range(S):
        K.wgrad(seg(d1, s), seg(zin, s), Gs[s][0], Bs, H, W, cout=HID, cin=Lp, g_s_co=c1 * 9, g_s_ci=9, ci_valid=L,
                defer=dr)
        with dr.side_section():
            K.plane_colsum(seg(d1, s), 0, HID, Bs, H, W, S=S1s[s], db=dbs[s][0])
            K.action_wgrad(S1s[s], a[s * Bs:(s + 1) * Bs], L, Gs[s][0])
    dz = torch.empty((B, L, H, W), dtype=torch.float32, device=dev)
    K.conv3x3(d1, wd[0], B, H, W, cin=HID, out_f32=dz, n_valid=L, dgrad=True, weights_stable=True)
    # spectral norm backward with the u, v currently held by the module (= last forward call)
    dr.join()
    dwbar = [torch.empty_like(w) if gw[i] is None else None for i, w in enumerate(wbar)]
    for s in range(S):
        K.spectral_norm_bwd([(Gs[s][i], wbar[i], u[i], v[i], sigma[0, i:i + 1], dots[s][i:i + 1],
                              dwbar[i] if gw[i] is None else gw[i], gw[i] is not None or s > 0,
                              sigma[s, i:i + 1]) for i in range(5)])
    if S == 1:
        db_out = [dbs[0][i].clone() if gb[i] is None else None for i in range(5)]
    else:  # bias gradients: sum_s colsum_s / rho_s
        inv_rho = (sigma / sigma[0:1])                                   # [S, 5] = 1 / rho_s
        tot = (torch.stack([torch.stack(dbs[s]) for s in range(S)]) * inv_rho.unsqueeze(-1)).sum(0)   # [5, HID]
        db_out = []
        for i in range(5):
            if gb[i] is None:
                db_out.append(tot[i].clone())
            else:
                gb[i] += tot[i]
                db_out.append(None)
    if gb6 is not None and Lp != L:
        gb6 += db6[:L]
    return (dz, dwbar, db_out, G6.clone() if gw6 is None else None, db6[:L].clone() if gb6 is None else None)


# ------------------------------------------------------------------------------------------------------------
# Encoder
# ------------------------------------------------------------------------------------------------------------
def encoder_forward(x, wbar, bias, sigma, w4, b4):
    """x: [B, 3C, H, W] fp32 view (dense CHW); sigma [3] from spectral_norm_update.  Returns (z, saved)."""
    dev = x.device
    B, Cin, H, W = x.shape
    Cp = _r16(Cin)
    L = w4.shape[0]
    Lp = _r16(L)
    wf = [K.packed_weight(HID, Cp, dev), K.packed_weight(HID, HID, dev), K.packed_weight(HID, HID, dev),
          K.packed_weight(Lp, HID, dev)]
    G16 = K.GRAD_DTYPE
    wd = [K.packed_weight(HID, HID, dev, G16), K.packed_weight(HID, HID, dev, G16), K.packed_weight(HID, Lp, dev, G16)]
    jobs = [_conv2d_fwd_job(wbar[i], wf[i], sigma[i:i + 1]) for i in range(3)] + [_conv2d_fwd_job(w4, wf[3])]
    jobs += [_conv2d_dgrad_job(wbar[1], wd[0], sigma[1:2]), _conv2d_dgrad_job(wbar[2], wd[1], sigma[2:3]),
             _conv2d_dgrad_job(w4, wd[2], None, co_valid=L)]
    K.pack_weights(jobs)
    xin = K.fwd_plane(B, H, W, Cp, dev)
    K.pack_nchw(x, xin, wrap=False)
    a1, a2, a3 = (K.fwd_plane(B, H, W, HID, dev) for _ in range(3))
    K.conv3x3(xin, wf[0], B, H, W, cin=Cp, bias=bias[0], act=ACT_LRELU, out=a1)
    K.conv3x3(a1, wf[1], B, H, W, cin=HID, bias=bias[1], act=ACT_LRELU, out=a2)
    K.conv3x3(a2, wf[2], B, H, W, cin=HID, bias=bias[2], act=ACT_LRELU, out=a3)
    z = torch.empty((B, L, H, W), dtype=torch.float32, device=dev)
    K.conv3x3(a3, wf[3], B, H, W, cin=HID, bias=b4, act=ACT_SIGMOID, out_f32=z, n_valid=L)
    return z, [xin, a1, a2, a3] + wd


# Called by encoder_backward with the indices (into [Wbar1..3, b1..3, W4, b4]) of the parameters whose gradient has just
# become final in stream order.  The data-parallel launcher hangs its per-layer all-reduce on it (scm_gan_b200/dp.py):
# the encoder's backward is the tail of every iteration, so its gradients are exchanged layer by layer while the layers
# below are still being differentiated.  Set through ops._encoder_backward; None = nobody listens.
LAYER_READY_HOOK = None


def _ready(idxs):
    if LAYER_READY_HOOK is not None:
        LAYER_READY_HOOK(idxs)


def encoder_backward(dz, z, saved, wbar, sigma, u, v, w4, sink=None):
    """sink = [gWbar1..3, gb1..3, gW4, gb4] (entries may be None): see transition_backward.
    Layer order: every layer's weight-gradient reduction runs on the side stream under the NEXT layer's data-gradient
    conv; it is joined, spectrally back-normalised and announced (LAYER_READY_HOOK) before the next weight gradient is
    issued, so that a data-parallel run can exchange it while the remaining layers are computed."""
    xin, a1, a2, a3 = saved[:4]
    wd = saved[4:]
    dev = dz.device
    B, L, H, W = dz.shape
    Lp = _r16(L)
    Cp = xin.shape[3]
    cin = wbar[0].shape[1]
    sizes = [w.numel() for w in wbar] + [w4.numel()]
    flat = torch.zeros(sum(sizes) + 3 * HID + Lp + 8, dtype=torch.float32, device=dev)
    G, off = [], 0
    for w, n in zip(list(wbar) + [w4], sizes):
        G.append(flat[off:off + n].view(w.shape))
        off += n
    db = [flat[off + i * HID: off + (i + 1) * HID] for i in range(3)]
    db4 = flat[off + 3 * HID: off + 3 * HID + Lp]
    dots = flat[off + 3 * HID + Lp:]
    sink = list(sink) if sink is not None else [None] * 8
    gw, gb, gw4, gb4 = sink[0:3], sink[3:6], sink[6], sink[7]
    if gw4 is not None:
        G[3] = gw4
    db = [db[i] if gb[i] is None else gb[i] for i in range(3)]
    if gb4 is not None and Lp == L:
        db4 = gb4
    dwbar = [torch.empty_like(w) if gw[i] is None else None for i, w in enumerate(wbar)]

    def finish_layer(i):  # spectral norm backward with the u, v currently held by the module (= last forward call)
        K.spectral_norm_bwd([(G[i], wbar[i], u[i], v[i], sigma[i:i + 1], dots[i:i + 1],
                              dwbar[i] if gw[i] is None else gw[i], gw[i] is not None)])
        _ready([i, 3 + i])

    dr = K.DeferredReduces(dev)
    d4 = K.new_plane(B, H, W, Lp, dev)
    K.pack_nchw(dz, d4, wrap=False, sig=z)
    K.wgrad(d4, a3, G[3], B, H, W, cout=Lp, cin=HID, g_s_co=HID * 9, g_s_ci=9, co_valid=L, defer=dr)
    with dr.side_section():
        K.plane_colsum(d4, 0, Lp, B, H, W, db=db4)
    d3, d2, d1 = (K.new_plane(B, H, W, HID, dev) for _ in range(3))
    K.conv3x3(d4, wd[2], B, H, W, cin=Lp, out=d3, gate=a3, dgrad=True)
    dr.join()                                                   # conv4: no spectral norm
    if gb4 is not None and Lp != L:
        gb4 += db4[:L]
    _ready([6, 7])
    K.wgrad(d3, a2, G[2], B, H, W, cout=HID, cin=HID, g_s_co=HID * 9, g_s_ci=9, db=db[2], defer=dr)
    K.conv3x3(d3, wd[1], B, H, W, cin=HID, out=d2, gate=a2, dgrad=True, weights_stable=True)
    dr.join()
    finish_layer(2)
    K.wgrad(d2, a1, G[1], B, H, W, cout=HID, cin=HID, g_s_co=HID * 9, g_s_ci=9, db=db[1], defer=dr)
    K.conv3x3(d2, wd[0], B, H, W, cin=HID, out=d1, gate=a1, dgrad=True, weights_stable=True)
    dr.join()
    finish_layer(1)
    K.wgrad(d1, xin, G[0], B, H, W, cout=HID, cin=Cp, g_s_co=cin * 9, g_s_ci=9, ci_valid=cin, db=db[0], defer=dr)
    dr.join()
    finish_layer(0)
    return (dwbar, [db[i].clone() if gb[i] is None else None for i in range(3)],
            G[3].clone() if gw4 is None else None, db4[:L].clone() if gb4 is None else None)


# ------------------------------------------------------------------------------------------------------------
# Decoder
# ------------------------------------------------------------------------------------------------------------
def decoder_forward(z, w1, b1, w2, b2):
    """z [B,L,H,W]; w1 ConvTranspose weight [L, 4L, 3, 3]; w2 [4L, Co, 3, 3] where Co is either the folded
    colour-channel count (latent groups pre-summed by the caller) or L*C for the visualisation path.
    Returns (logits [B,Co,H,W], saved)."""
    dev = z.device
    B, L, H, W = z.shape
    Lp = _r16(L)
    hid = w1.shape[1]
    co = w2.shape[1]
    cop = _r16(co)
    assert hid % 64 == 0 and hid <= HID
    # the hidden plane is 128 channels wide (wgrad wants a 128-channel operand) but only `hid` are computed: the
    # unwritten upper channels only ever reach wgrad outputs beyond co_valid / ci_valid, which are discarded
    wf1 = K.packed_weight(hid, Lp, dev)
    wf2 = K.packed_weight(cop, hid, dev)
    wd1 = K.packed_weight(Lp, hid, dev, K.GRAD_DTYPE)
    wd2 = K.packed_weight(hid, cop, dev, K.GRAD_DTYPE)
    K.pack_weights([_convT_fwd_job(w1, wf1), _convT_fwd_job(w2, wf2), _convT_dgrad_job(w1, wd1),
                    _convT_dgrad_job(w2, wd2)])
    zin = K.fwd_plane(B, H, W, Lp, dev)
    K.pack_nchw(z, zin, wrap=False)
    hidp = K.fwd_plane(B, H, W, HID, dev)
    K.conv3x3(zin, wf1, B, H, W, cin=Lp, bias=b1, act=ACT_LRELU, out=hidp)
    logits = torch.empty((B, co, H, W), dtype=torch.float32, device=dev)
    K.conv3x3(hidp, wf2, B, H, W, cin=hid, bias=b2, act=ACT_NONE, out_f32=logits, n_valid=co)
    return logits, [zin, hidp, wd1, wd2]


def decoder_bce_forward(z, w1, b1, w2, b2, target_bt, mask_bt):
    """Decoder + pixel loss of all T rollout steps (reference main.py:188-197 once per step): z [T*B, L, H, W] t-major,
    w2 already folded over the latent groups (Co = colour channels <= 16), target_bt [B, T, C, H, W], mask_bt [B, T].
    The last conv's epilogue evaluates sigmoid + BCE + the masked means (scmgan_decoder_bce_fwd): no logits tensor, no
    separate loss kernel, and d loss_t / d logits arrives as the gradient plane the backward convolutions read.
    Returns (loss_t [T], saved = [zin, hidp, wd1, wd2, d2])."""
    dev = z.device
    TB, L, H, W = z.shape
    B, T = target_bt.shape[0], target_bt.shape[1]
    assert TB == T * B
    Lp = _r16(L)
    hid = w1.shape[1]
    co = w2.shape[1]
    cop = _r16(co)
    assert hid % 64 == 0 and hid <= HID and cop == 16 and target_bt.shape[2] == co
    wf1 = K.packed_weight(hid, Lp, dev)
    wf2 = K.packed_weight(cop, hid, dev)
    wd1 = K.packed_weight(Lp, hid, dev, K.GRAD_DTYPE)
    wd2 = K.packed_weight(hid, cop, dev, K.GRAD_DTYPE)
    K.pack_weights([_convT_fwd_job(w1, wf1), _convT_fwd_job(w2, wf2), _convT_dgrad_job(w1, wd1),
                    _convT_dgrad_job(w2, wd2)])
    zin = K.fwd_plane(TB, H, W, Lp, dev)
    K.pack_nchw(z, zin, wrap=False)
    hidp = K.fwd_plane(TB, H, W, HID, dev)
    K.conv3x3(zin, wf1, TB, H, W, cin=Lp, bias=b1, act=ACT_LRELU, out=hidp)
    loss_t = torch.zeros(T, dtype=torch.float32, device=dev)
    d2 = K.new_plane(TB, H, W, cop, dev)
    K.decoder_bce_fwd(hidp, wf2, T, B, H, W, cin=hid, bias=b2, n_valid=co, dlogits_plane=d2, target_bt=target_bt,
                      mask_bt=mask_bt, loss_t=loss_t)
    return loss_t, [zin, hidp, wd1, wd2, d2]


def decoder_bce_backward(g, saved, w1, w2, T, sink=None):
    """Backward of decoder_bce_forward.  g [T] = d total / d loss_t (the gradient plane is rescaled in place only where
    g[t] != 1).  Returns like decoder_backward."""
    zin, hidp, wd1, wd2, d2 = saved
    TB, Hp, Wp, _ = d2.shape
    K.decoder_bce_bwd(d2, g, T, TB // T, Hp - 2, Wp - 2)
    return decoder_backward(None, [zin, hidp, wd1, wd2], w1, w2, sink, d2=d2)


def decoder_backward(dlogits, saved, w1, w2, sink=None, d2=None):
    """sink = [gW1, gb1, gW2, gb2] (entries may be None): see transition_backward.
    d2 (optional): d loss / d logits already as a gradient plane (fused loss head); dlogits is then ignored."""
    zin, hidp, wd1, wd2 = saved
    dev = zin.device
    co = w2.shape[1]
    B, H, W = zin.shape[0], zin.shape[1] - 2, zin.shape[2] - 2
    cop = _r16(co)
    L, hid = w1.shape[0], w1.shape[1]
    Lp = zin.shape[3]
    flat = torch.zeros(w1.numel() + w2.numel() + HID + cop, dtype=torch.float32, device=dev)
    g1 = flat[:w1.numel()].view(w1.shape)
    g2 = flat[w1.numel(): w1.numel() + w2.numel()].view(w2.shape)
    db1 = flat[w1.numel() + w2.numel(): w1.numel() + w2.numel() + HID]
    db2 = flat[w1.numel() + w2.numel() + HID:]
    sink = list(sink) if sink is not None else [None] * 4
    if sink[0] is not None:
        g1 = sink[0]
    if sink[1] is not None and hid % 8 == 0:   # wgrad writes the valid channels rounded up to 8
        db1 = sink[1]
    if sink[2] is not None:
        g2 = sink[2]
    if sink[3] is not None and cop == co:
        db2 = sink[3]
    dr = K.DeferredReduces(dev)
    if d2 is None:
        d2 = K.new_plane(B, H, W, cop, dev)
        K.pack_nchw(dlogits, d2, wrap=False)
    # ConvTranspose weight layout [Cin][Cout][3][3], taps flipped relative to the equivalent correlation
    K.wgrad(d2, hidp, g2, B, H, W, cout=cop, cin=HID, g_s_co=9, g_s_ci=co * 9, flip=True, co_valid=co, ci_valid=hid, defer=dr)
    with dr.side_section():
        K.plane_colsum(d2, 0, cop, B, H, W, db=db2)
    d1 = K.new_plane(B, H, W, HID, dev)
    K.conv3x3(d2, wd2, B, H, W, cin=cop, out=d1, gate=hidp, dgrad=True)
    K.wgrad(d1, zin, g1, B, H, W, cout=HID, cin=Lp, g_s_co=9, g_s_ci=hid * 9, flip=True, co_valid=hid, ci_valid=L,
            db=db1, defer=dr)
    dz = torch.empty((B, L, H, W), dtype=torch.float32, device=dev)
    K.conv3x3(d1, wd1, B, H, W, cin=hid, out_f32=dz, n_valid=L, dgrad=True, weights_stable=True)
    dr.join()
    if sink[1] is not None and db1 is not sink[1]:
        sink[1].add_(db1[:hid])
    if sink[3] is not None and db2 is not sink[3]:
        sink[3].add_(db2[:co])
    return (dz, g1.clone() if sink[0] is None else None, db1[:hid].clone() if sink[1] is None else None,
            g2.clone() if sink[2] is None else None, db2[:co].clone() if sink[3] is None else None)


# ------------------------------------------------------------------------------------------------------------
# RewardPredictor (reference models.py:235-250)
# ------------------------------------------------------------------------------------------------------------
RHID = 32  # hidden width of the reward head (models.py:230)


def reward_forward(z, w1, b1, w2, b2, want_map):
    """conv(L->32, valid) -> LeakyReLU -> conv(32->3R, stride 2, valid) -> softmax(3) -> p+ - p- -> spatial sum.
    Both convs run as same-size stride-1 tcgen05 convs; valid/stride-2 positions are selected by the head kernel
    (the footprint of every consumed output lies inside the valid region, so the results are identical)."""
    dev = z.device
    B, L, H, W = z.shape
    Lp = _r16(L)
    co = w2.shape[0]
    R = co // 3
    assert w1.shape[0] == RHID and co <= 16
    # hidden plane 128 channels wide for wgrad, only RHID computed (see decoder_forward)
    # the hidden layer is computed 64 channels wide (32 real + 32 zero weight rows): the 16-channel-input kernel
    # (conv_expand.cuh) stores 64-channel halves
    wf1 = K.packed_weight(2 * RHID, Lp, dev)
    wf2 = K.packed_weight(16, RHID, dev)
    wd1 = K.packed_weight(Lp, RHID, dev, K.GRAD_DTYPE)
    wd2 = K.packed_weight(2 * RHID, 16, dev, K.GRAD_DTYPE)
    K.pack_weights([_conv2d_fwd_job(w1, wf1), _conv2d_fwd_job(w2, wf2),
                    _conv2d_dgrad_job(w1, wd1), _conv2d_dgrad_job(w2, wd2)])
    zin = K.fwd_plane(B, H, W, Lp, dev)
    K.pack_nchw(z, zin, wrap=False)
    hidp = K.fwd_plane(B, H, W, HID, dev)
    K.conv3x3(zin, wf1, B, H, W, cin=Lp, bias=b1, act=ACT_LRELU, out=hidp)
    y2 = torch.empty((B, co, H, W), dtype=torch.float32, device=dev)
    K.conv3x3(hidp, wf2, B, H, W, cin=RHID, bias=b2, act=ACT_NONE,
              out_f32=y2, n_valid=co)
    r = torch.empty((B, R), dtype=torch.float32, device=dev)
    h2, w2s = (H - 5) // 2 + 1, (W - 5) // 2 + 1
    rmap = torch.empty((B, R, h2, w2s), dtype=torch.float32, device=dev) if want_map else None
    K.reward_head_fwd(y2, R, r, rmap)
    return r, rmap, [zin, hidp, y2, wd1, wd2]


def reward_backward(dr, saved, w1, w2, sink=None):
    """sink = [gW1, gb1, gW2, gb2] (entries may be None): see transition_backward."""
    zin, hidp, y2, wd1, wd2 = saved
    dev = dr.device
    B, co, H, W = y2.shape
    R = co // 3
    L = w1.shape[1]
    Lp = zin.shape[3]
    n1, n2 = w1.numel(), w2.numel()
    flat = torch.zeros(n1 + n2 + HID + 16, dtype=torch.float32, device=dev)
    g1, g2 = flat[:n1].view(w1.shape), flat[n1:n1 + n2].view(w2.shape)
    db1, db2 = flat[n1 + n2:n1 + n2 + HID], flat[n1 + n2 + HID:]
    sink = list(sink) if sink is not None else [None] * 4
    if sink[0] is not None:
        g1 = sink[0]
    if sink[1] is not None:   # RHID = 32 valid channels are written
        db1 = sink[1]
    if sink[2] is not None:
        g2 = sink[2]
    red = K.DeferredReduces(dev)
    d2 = K.new_plane(B, H, W, 16, dev)
    K.reward_head_bwd(y2, dr, R, d2)
    K.wgrad(d2, hidp, g2, B, H, W, cout=16, cin=HID, g_s_co=RHID * 9, g_s_ci=9, co_valid=co, ci_valid=RHID, defer=red)
    with red.side_section():
        K.plane_colsum(d2, 0, 16, B, H, W, db=db2)
    d1 = K.new_plane(B, H, W, HID, dev)
    K.conv3x3(d2, wd2, B, H, W, cin=16, out=d1, gate=hidp, dgrad=True)
    K.wgrad(d1, zin, g1, B, H, W, cout=HID, cin=Lp, g_s_co=L * 9, g_s_ci=9, co_valid=RHID, ci_valid=L, db=db1, defer=red)
    dz = torch.empty((B, L, H, W), dtype=torch.float32, device=dev)
    K.conv3x3(d1, wd1, B, H, W, cin=RHID, out_f32=dz, n_valid=L, dgrad=True, weights_stable=True)
    red.join()
    if sink[3] is not None:
        sink[3].add_(db2[:co])
    return (dz, g1.clone() if sink[0] is None else None, db1[:RHID].clone() if sink[1] is None else None,
            g2.clone() if sink[2] is None else None, db2[:co].clone() if sink[3] is None else None)


# ------------------------------------------------------------------------------------------------------------
# CoordConv2d with a 3x3 / stride 1 / zero-pad 1 convolution (reference coordconv.py:5-15; interface-only layer)
# ------------------------------------------------------------------------------------------------------------
def coordconv_forward(x, w, b):
    """x [B,C,H,W] fp32; w [Co, C+2, 3, 3]; b [Co] or None.  The two coordinate channels are synthesised straight
    into the bf16 input plane (never concatenated in fp32).  Returns (y [B,Co,H,W], saved)."""
    dev = x.device
    B, Cc, H, W = x.shape
    co, cin = w.shape[0], w.shape[1]
    assert cin == Cc + 2 and Cc % 2 == 0, "CoordConv2d expects in_channels + 2 (even number of data channels)"
    Cp, Np = _r16(cin), _r16(co)
    wf = K.packed_weight(Np, Cp, dev)
    wd = K.packed_weight(_r16(Cc), HID if Np <= HID else Np, dev, K.GRAD_DTYPE)  # dgrad reads the 128-wide gradient plane
    K.pack_weights([_conv2d_fwd_job(w, wf), _conv2d_dgrad_job(w, wd, None, 0, Cc)])
    xin = K.fwd_plane(B, H, W, Cp, dev)
    K.pack_nchw(x, xin, c_pad=Cp, wrap=False)
    y = torch.empty((B, co, H, W), dtype=torch.float32, device=dev)
    if coords_in_tile(Cp, W):
        # the coordinate channels are generated by the conv's producer while it stages the im2col tile; the plane
        # holds zeros in their place (scmgan_conv_desc::coord_c1)
        K.conv3x3(xin, wf, B, H, W, cin=Cp, bias=b, act=ACT_NONE, out_f32=y, n_valid=co, coord_c=Cc)
    else:
        K.pack_coords(xin, Cc)   # wide layers (TMA-fed tiles): coordinates materialised in the 16-bit input plane
        K.conv3x3(xin, wf, B, H, W, cin=Cp, bias=b, act=ACT_NONE, out_f32=y, n_valid=co)
    return y, [xin, wd]


def coords_in_tile(Cp, W):
    """In-tile coordinate generation is available where the conv's producer stages the tile in software: 16 input
    channels (data + 2 coordinates) and W <= 69."""
    return Cp == 16 and W + 2 <= 71


def coordconv_backward(dy, saved, w):
    """-> (dx [B,C,H,W], dW [Co,C+2,3,3], db [Co]).  The gradient plane is 128 channels wide (zero beyond Co) because
    the weight-gradient kernels want a 128-channel operand."""
    xin, wd = saved
    dev = dy.device
    B, co, H, W = dy.shape
    cin = w.shape[1]
    Cc = cin - 2
    Cp = xin.shape[3]
    Gp = wd.shape[2]
    assert co <= Gp
    dyp = K.new_plane(B, H, W, Gp, dev)
    K.pack_nchw(dy, dyp, c_pad=Gp, wrap=False)
    g = torch.zeros_like(w)
    db = torch.zeros(Gp, dtype=torch.float32, device=dev)
    if coords_in_tile(Cp, W):
        # data channels from the plane (its coordinate channels are zero); the coordinate channels' gradient is the
        # correlation of dy with the two ramps (scmgan_coord_wgrad)
        K.wgrad(dyp, xin, g, B, H, W, cout=Gp, cin=Cp, g_s_co=cin * 9, g_s_ci=9, co_valid=co, ci_valid=Cc, db=db)
        K.coord_wgrad(dy.contiguous(), g, Cc)
    else:
        K.wgrad(dyp, xin, g, B, H, W, cout=Gp, cin=Cp, g_s_co=cin * 9, g_s_ci=9, co_valid=co, ci_valid=cin, db=db)
    dx = torch.empty((B, Cc, H, W), dtype=torch.float32, device=dev)
    K.conv3x3(dyp, wd, B, H, W, cin=Gp, out_f32=dx, n_valid=Cc, dgrad=True)
    return dx, g, db[:co].clone()
