"""TEST INFRASTRUCTURE ONLY - never imported by the product path (scm_gan_b200/).

CPU/fp32 restatement of the reference's world-model training path as pure functions over reference-format
state_dicts.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.

Parity status: the reference ships no tests, seeds or golden vectors, so this oracle is
pinned against outputs of the reference's own modules executed in the build container
(oracle/make_golden.py -> tests/golden/*.pt, checked by tests/test_oracle_golden.py and, when /root/reference
is present, directly against the live reference classes by tests/test_oracle_vs_reference.py).

Every function cites the reference lines it restates.  All arithmetic is torch fp32; the padding semantics are
the legacy (torch 1.1-1.4) circular pad-1 under which the reference ran (oracle/shims.py, shim 2).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

ENCODER_INPUT_FRAMES = 3  # reference models.py:19
CF_REGULARIZATION_RATE = 5  # reference main.py:54
CF_REGULARIZATION_LAMBDA = 0.01  # reference main.py:55


# --------------------------------------------------------------------------------------------------------------
# spectral_normalization.py
# --------------------------------------------------------------------------------------------------------------
def l2normalize(v, eps=1e-12):
    """reference spectral_normalization.py:10-11"""
    return v / (v.norm() + eps)


def spectral_norm_weight(sd, prefix):
    """reference spectral_normalization.py:23-35 (`_update_u_v`, power_iterations=1).

    u and v are advanced through `.data` assignments exactly like the reference.  Consequence (verified against
    the live reference, see DESIGN.md "SpectralNorm backward"): autograd saves the u/v *Parameters*, not their
    values, so when a wrapped conv is called several times before backward() (every rollout step), the gradient
    of every call's sigma is taken with the u, v of the LAST call:
        dWbar = sum_t [ G_t/sigma_t - (<G_t, Wbar>/sigma_t^2) * u_last v_last^T ].
    Returns the normalised weight (differentiable w.r.t. weight_bar through the division and sigma)."""
    u, v, w = sd[prefix + "weight_u"], sd[prefix + "weight_v"], sd[prefix + "weight_bar"]
    height = w.shape[0]
    v.data = l2normalize(torch.mv(torch.t(w.view(height, -1).data), u.data))
    u.data = l2normalize(torch.mv(w.view(height, -1).data, v.data))
    sigma = u.dot(w.view(height, -1).mv(v))
    return w / sigma.expand_as(w)


# --------------------------------------------------------------------------------------------------------------
# Optional operand rounding ("bf16-operand restatement").  With OPERAND_DTYPE = None (default) everything below is
# the plain fp32 algorithm.  With OPERAND_DTYPE = torch.bfloat16 the *same* algorithm is evaluated with conv inputs,
# normalised weights and the gradients entering each conv's backward rounded to that dtype (fp32 accumulation), i.e.
# the arithmetic contract BASELINE.json's north_star prescribes for the tensor-core path.  Tests use it to separate
# "kernel bug" from "operand precision": LeakyReLU kinks turn a 2^-9 operand rounding into percent-level L2
# differences of fp32 gradients (sign flips of near-zero pre-activations), see DESIGN.md "Parity".
# --------------------------------------------------------------------------------------------------------------
OPERAND_DTYPE = None
# Finer control for the precision study (profiles/grad_precision.py): rounding of the forward activations, of the
# (normalised) weights and of the gradients entering each conv's backward can be chosen separately; each defaults to
# OPERAND_DTYPE.  Values: None (fp32), a torch dtype, or "tf32" (10-bit mantissa, fp32 range, round-to-nearest-even
# emulated on the fp32 bit pattern).
ROUND = {"act": "inherit", "w": "inherit", "grad": "inherit"}
# When a list, every LeakyReLU appends the sign pattern (pre-activation > 0) of its input: lets a study count how many
# kinks two evaluations of the same network put on different sides.
RECORD_SIGNS = None


def _dtype_of(kind):
    d = ROUND[kind]
    return OPERAND_DTYPE if d == "inherit" else d


def _round_to(x, dtype):
    if dtype == "tf32":
        i = x.contiguous().view(torch.int32)
        i = (i + 0x0FFF + ((i >> 13) & 1)) & ~0x1FFF  # round to nearest even at 13 dropped mantissa bits
        return i.view(torch.float32).view_as(x)
    return x.to(dtype).to(x.dtype)


class _RoundSTE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, dtype):
        return _round_to(x, dtype)

    @staticmethod
    def backward(ctx, g):
        return g, None


class _RoundGrad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, dtype):
        ctx.dtype = dtype
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return _round_to(g, ctx.dtype), None


def _q(x):
    """activation-operand rounding (forward), straight-through gradient"""
    d = _dtype_of("act")
    return x if d is None else _RoundSTE.apply(x, d)


def _qw(w):
    """weight-operand rounding (forward), straight-through gradient"""
    d = _dtype_of("w")
    return w if d is None else _RoundSTE.apply(w, d)


def _qg(x):
    """identity forward; the gradient flowing back through this point is rounded (gradient planes are bf16)"""
    d = _dtype_of("grad")
    return x if d is None else _RoundGrad.apply(x, d)


def _lrelu(x, tag=""):
    """F.leaky_relu with the default slope 0.01 (reference models.py:77-99, 145-151, 239, 276)"""
    if RECORD_SIGNS is not None:
        RECORD_SIGNS.append((tag, x.detach() > 0))
    return F.leaky_relu(x)


def _rounding_on():
    return any(_dtype_of(k) is not None for k in ("act", "w", "grad"))


def _sn_conv(sd, name, x, circular, exact_tail=0):
    """SpectralNorm(nn.Conv2d(..., 3x3, stride 1)) forward: reference spectral_normalization.py:66-68 +
    models.py:51-55 / 129-133.  circular => legacy wrap-by-one padding, else zero padding 1.
    exact_tail: number of trailing input channels kept in fp32 under operand rounding (the spatially constant
    action channels of Transition.conv1, which the product folds into an fp32 per-sample bias)."""
    w = spectral_norm_weight(sd, name + ".module.")
    b = sd[name + ".module.bias"]
    if _rounding_on() and exact_tail:
        n = x.shape[1] - exact_tail
        xq = torch.cat([_q(x[:, :n]), x[:, n:]], dim=1)
        wq = torch.cat([_qw(w[:, :n]), w[:, n:]], dim=1)
    else:
        xq, wq = _q(x), _qw(w)
    if circular:
        return _qg(F.conv2d(F.pad(xq, (1, 1, 1, 1), mode="circular"), wq, b))
    return _qg(F.conv2d(xq, wq, b, padding=1))


# --------------------------------------------------------------------------------------------------------------
# models.py
# --------------------------------------------------------------------------------------------------------------
def encoder_forward(sd, x):
    """reference models.py:139-157.  x: [B, 3, C, H, W] -> [B, latent, H, W].  bn_conv1 is registered
    (models.py:130) but never applied."""
    b, frames, ch, h, w = x.shape
    x = x.reshape(b, frames * ch, h, w)
    x = _lrelu(_sn_conv(sd, "conv1", x, False), "enc.conv1")
    x = _lrelu(_sn_conv(sd, "conv2", x, False), "enc.conv2")
    x = _lrelu(_sn_conv(sd, "conv3", x, False), "enc.conv3")
    x = _qg(F.conv2d(_q(x), _qw(sd["conv4.weight"]), sd["conv4.bias"], padding=1))
    return torch.sigmoid(x)


class _StraightThroughBernoulli(torch.autograd.Function):
    """reference models.py:30-40 with the sample drawn as (U < p) from injected uniforms (shims.py, shim 5)."""

    @staticmethod
    def forward(ctx, p, u):
        return (u < p).to(p.dtype)

    @staticmethod
    def backward(ctx, g):
        return g, None


def transition_forward(sd, z, a, training=True, uniforms=None, return_all=False, return_probs=False):
    """reference models.py:59-119.  z: [B, L, H, W]; a: [B, A] (one-hot); uniforms: [B, L, H, W] when training."""
    b, _, h, w = z.shape
    assert a.shape[0] == b  # models.py:66
    actions = a.unsqueeze(-1).unsqueeze(-1).repeat(1, 1, h, w)
    x = torch.cat([z, actions], dim=1)
    x = _lrelu(_sn_conv(sd, "conv1", x, True, exact_tail=a.shape[1]), "tr.conv1")
    skip1 = x
    x = _lrelu(_sn_conv(sd, "conv2", x, True), "tr.conv2")
    skip2 = x
    x = _lrelu(_sn_conv(sd, "conv3", x, True), "tr.conv3")
    out3 = x
    x = _lrelu(_sn_conv(sd, "conv4", x, True), "tr.conv4")
    out4 = x
    x = torch.cat([x, skip2], dim=1)
    x = _lrelu(_sn_conv(sd, "conv5", x, True), "tr.conv5")
    out5 = x
    x = torch.cat([x, skip1], dim=1)
    x = _qg(F.conv2d(F.pad(_q(x), (1, 1, 1, 1), mode="circular"), _qw(sd["conv6.weight"]), sd["conv6.bias"]))
    p = torch.sigmoid(x)
    if training:
        if callable(uniforms):  # hook(p) -> uniforms, lets a test keep its draws away from p (margin-safe sampling)
            uniforms = uniforms(p.detach())
        if uniforms is None:
            uniforms = torch.rand_like(p)
        x = _StraightThroughBernoulli.apply(p, uniforms)
    else:
        x = (p > 0.5).to(p.dtype)
    if return_all:
        return (skip1, skip2, out3, out4, out5, x)
    if return_probs:
        return x, p
    return x


def decoder_forward(sd, z, visualize=False):
    """reference models.py:270-291.  Returns logits [B, C, H, W] (caller applies sigmoid, main.py:189)."""
    b, latent, h, w = z.shape
    x = _qg(F.conv_transpose2d(_q(z), _qw(sd["conv1.weight"]), sd["conv1.bias"], stride=1, padding=1))
    x = _lrelu(x, "dec.conv1")
    w2, b2 = sd["conv2.weight"], sd["conv2.bias"]
    color = w2.shape[1] // latent
    if _rounding_on() and not visualize:
        # the product sums the latent groups of the (linear) last layer in fp32 *before* rounding the weights
        w2f = w2.view(w2.shape[0], latent, color, 3, 3).sum(1)
        b2f = b2.view(latent, color).sum(0)
        return _qg(F.conv_transpose2d(_q(x), _qw(w2f), b2f, stride=1, padding=1))
    x = _qg(F.conv_transpose2d(_q(x), _qw(w2), b2, stride=1, padding=1))
    x = x.view(b, latent, color, h, w)
    vis = x[0]
    x = torch.sum(x, dim=1)
    if visualize:
        return x, vis
    return x


def reward_forward(sd, z, visualize=False):
    """reference models.py:235-250.  [B, L, H, W] -> [B, R]"""
    x = _lrelu(_qg(F.conv2d(_q(z), _qw(sd["conv1.weight"]), sd["conv1.bias"])), "rew.conv1")
    x = _qg(F.conv2d(_q(x), _qw(sd["conv2.weight"]), sd["conv2.bias"], stride=2))
    b, ch, h, w = x.shape
    x = torch.softmax(x.view(b, 3, ch // 3, h, w), dim=1)
    x = x[:, 0] - x[:, 2]
    if visualize:
        return x.sum(-1).sum(-1), x
    return x.sum(-1).sum(-1)


def coordconv_forward(weight, bias, x, **conv_kwargs):
    """reference coordconv.py:10-15 (square inputs only - the reference builds coord_x as (W, W))."""
    b, _, h, w = x.shape
    cx = torch.arange(-1.0, 1.0, 2 / w).unsqueeze(0).repeat(w, 1).unsqueeze(0).repeat(b, 1, 1).unsqueeze(1)
    cy = torch.arange(-1.0, 1.0, 2 / h).unsqueeze(1).repeat(1, h).unsqueeze(0).repeat(b, 1, 1).unsqueeze(1)
    x = torch.cat([x, cx.to(x), cy.to(x)], dim=1)
    return F.conv2d(x, weight, bias, **conv_kwargs)


def _gru_step(x, hprev, w_ih, w_hh):
    """single-step nn.GRU(bias=False), gate order r, z, n (torch docs; reference spatial_recurrent.py:31-34)."""
    gi = x @ w_ih.t()
    gh = hprev @ w_hh.t()
    i_r, i_z, i_n = gi.chunk(3, 1)
    h_r, h_z, h_n = gh.chunk(3, 1)
    r = torch.sigmoid(i_r + h_r)
    zg = torch.sigmoid(i_z + h_z)
    n = torch.tanh(i_n + r * h_n)
    return (1 - zg) * n + zg * hprev


def csrn_forward(sd, x):
    """reference spatial_recurrent.py:46-119, including its quirks: the right sweep writes into context_left
    (line 110), so context_right stays zero and the left sweep's results are overwritten."""
    b, c, h, w = x.shape
    ctx = {k: torch.zeros(b, c, h, w, dtype=x.dtype, device=x.device) for k in ("above", "below", "left", "right")}

    def sweep(direction, rnn, conv, target, indices, vertical):
        n = w if vertical else h
        state = torch.zeros(b * n, c, dtype=x.dtype, device=x.device)
        for i in indices:
            line = x[:, :, i, :] if vertical else x[:, :, :, i]
            line = line.permute(0, 2, 1).contiguous().view(b * n, c)
            out = _gru_step(line, state, sd[rnn + ".weight_ih_l0"], sd[rnn + ".weight_hh_l0"])
            conv_in = out.view(b, n, c).permute(0, 2, 1)
            if vertical:
                ctx[target][:, :, i, :] = conv_in
            else:
                ctx[target][:, :, :, i] = conv_in
            conv_out = torch.tanh(F.conv1d(conv_in, sd[conv + ".weight"], sd[conv + ".bias"], padding=1))
            state = conv_out.permute(0, 2, 1).contiguous().view(b * n, c)

    sweep("down", "rnn_down", "conv_down", "above", range(h), True)
    sweep("up", "rnn_up", "conv_up", "below", reversed(range(h)), True)
    sweep("left", "rnn_left", "conv_left", "left", range(w), False)
    sweep("right", "rnn_right", "conv_right", "left", reversed(range(w)), False)
    cmap = torch.cat((ctx["above"], ctx["below"], ctx["left"], ctx["right"]), dim=1)
    return F.conv2d(cmap, sd["conv_combine.weight"], sd["conv_combine.bias"])


# --------------------------------------------------------------------------------------------------------------
# main.py
# --------------------------------------------------------------------------------------------------------------
def decoder_pixel_loss(target, predicted):
    """reference main.py:310-312"""
    return F.binary_cross_entropy(predicted, target, reduction="none").mean(-1).mean(-1).mean(-1)


def latent_state_loss(target, predicted):
    """reference main.py:306-307"""
    return ((target - predicted) ** 2).mean(-1).mean(-1).mean(-1)


def prediction_horizon(train_iter, train_iters, hmin=3, hmax=10):
    """reference main.py:143-145"""
    theta = train_iter / train_iters
    return hmin + int((hmax - hmin) * theta), theta


def train_step_loss(nets, states, rewards, dones, actions, *, num_actions, theta, reward_coef=1e-3,
                    uniforms=None, truncate_bptt=False, enable_disentanglement=False, enable_action_control=False,
                    cf_now=False, counterfactual_horizon=1, cf_indices=None, cf_perm=None, latent_dim=16,
                    latent_overshooting=False, td_lambda=0.9):
    """reference main.py:155-283 (loss construction of one training iteration, incl. the optional latent
    overshooting of main.py:217-234).

    nets: dict with reference-format state_dicts 'encoder', 'transition', 'decoder', 'reward_predictor'
          (tensors with requires_grad where a gradient is wanted; SN u/v are advanced in place).
    states [B,Hn,C,H,W], rewards [B,Hn,R], dones [B,Hn] (float), actions [B,Hn] (int64 numpy or tensor).
    uniforms: list of [B,L,H,W] tensors, one per Transition call in call order (shim 5), or None.
    cf_indices: [B,2] ints (idx_a, idx_b) replacing np.random.randint (main.py:249-250);
    cf_perm: batch permutation replacing np.random.shuffle(cf_actions) (main.py:275).
    Returns (loss, dict of named loss terms, final z).
    """
    # a device tensor stays where it is (graph-captured stock-torch baseline); numpy / lists as in main.py:206
    actions = actions.long() if torch.is_tensor(actions) else torch.as_tensor(np.asarray(actions)).long()
    bsz = states.shape[0]
    hn = states.shape[1]
    eye = torch.eye(num_actions, dtype=states.dtype, device=states.device)
    it = iter(uniforms) if (uniforms is not None and not callable(uniforms)) else None

    def trans(zz, aa):
        u = next(it) if it is not None else uniforms
        return transition_forward(nets["transition"], zz, aa, training=True, uniforms=u)

    z = encoder_forward(nets["encoder"], states[:, 0:3])  # main.py:162
    z_orig = z.clone()  # main.py:163
    active_mask = torch.ones(bsz, dtype=states.dtype, device=states.device)
    loss = 0
    lo_loss = 0
    lo_z_set = {}
    terms = {}
    for t in range(1, hn - 1):  # main.py:177
        active_mask = active_mask * (1 - dones[:, t])
        expected_reward = reward_forward(nets["reward_predictor"], z)
        reward_difference = torch.mean(torch.mean((expected_reward - rewards[:, t]) ** 2, dim=1) * active_mask)
        terms[f"Rd Loss t={t}"] = reward_difference
        loss = loss + theta * reward_coef * reward_difference
        predicted = torch.sigmoid(decoder_forward(nets["decoder"], z))
        rec_loss_batch = decoder_pixel_loss(states[:, t], predicted)
        if truncate_bptt and t > 1:
            z = z.detach()  # main.py:192-193 (detach_ on the loop variable)
        rec_loss = torch.mean(rec_loss_batch * active_mask)
        terms[f"Reconstruction t={t}"] = rec_loss
        loss = loss + rec_loss
        z = trans(z, eye[actions[:, t]])  # main.py:206-207

        if latent_overshooting:  # main.py:217-230
            lo_z_set[t] = encoder_forward(nets["encoder"], states[:, t - 1:t + 2])
            for t_left in range(1, t):
                lo_z_set[t_left] = trans(lo_z_set[t_left], eye[actions[:, t - 1]])
            for t_a in range(2, t - 1):
                lo_loss_batch = latent_state_loss(lo_z_set[t].detach(), lo_z_set[t_a])
                lo_loss = lo_loss + td_lambda * torch.mean(lo_loss_batch * active_mask)

    if latent_overshooting:  # main.py:232-234
        terms["LO total"] = lo_loss
        loss = loss + theta * lo_loss

    if enable_disentanglement and cf_now:  # main.py:242-262
        z_cf_a = z.clone()
        z_cf_b = z_orig
        unswapped = torch.ones((bsz, latent_dim), dtype=states.dtype, device=states.device)
        for i in range(bsz):
            idx_a, idx_b = int(cf_indices[i][0]), int(cf_indices[i][1])
            unswapped[i, idx_a].zero_()   # (= 0; written as a device-side fill so that a CUDA-graph capture of the
            unswapped[i, idx_b].zero_()   #  stock-torch baseline does not need a host scalar upload)
            # main.py:253: tuple assignment on views => net effect z[i,idx_a] <- z[i,idx_b]
            z_cf_b[i, idx_a], z_cf_b[i, idx_b] = z_cf_b[i, idx_b], z_cf_b[i, idx_a]
        for t in range(1, counterfactual_horizon):
            z_cf_b = trans(z_cf_b, eye[actions[:, t]])
        cf_loss = torch.abs(z_cf_a - z_cf_b).mean(-1).mean(-1) * unswapped
        cf_loss = CF_REGULARIZATION_LAMBDA * torch.mean(cf_loss.mean(-1) * active_mask)
        loss = loss + cf_loss
        terms["CF Disentanglement Loss"] = cf_loss

    if enable_action_control and cf_now:  # main.py:268-283
        z_cf_a = z.clone()
        z_cf_b = z_orig
        cf_actions = actions[cf_perm.long() if torch.is_tensor(cf_perm) else torch.as_tensor(np.asarray(cf_perm)).long()]
        for t in range(1, counterfactual_horizon):
            z_cf_b = trans(z_cf_b, eye[cf_actions[:, t]])
        eps = 0.001
        cf_loss = -torch.log(torch.abs(z_cf_a - z_cf_b).mean(-1).mean(-1).mean(-1) + eps)
        cf_loss = CF_REGULARIZATION_LAMBDA * torch.mean(cf_loss * active_mask)
        loss = loss + cf_loss
        terms["CF Control Bias Loss"] = cf_loss
    return loss, terms, z


def measure_prediction_mse(nets, states, rewards, dones, actions, *, num_actions):
    """reference main.py:784-836 (eval-mode rollout MSE; returns the per-step lists instead of plotting)."""
    actions = torch.as_tensor(np.asarray(actions)).long()
    bsz, timesteps = states.shape[0], states.shape[1]
    eye = torch.eye(num_actions, dtype=states.dtype, device=states.device)
    with torch.no_grad():
        z = encoder_forward(nets["encoder"], states[:, :3])
        z = transition_forward(nets["transition"], z, eye[actions[:, 1]], training=False)
        mse, mse_std, rew, rew_std = [], [], [], []
        active_mask = torch.ones(bsz, dtype=states.dtype, device=states.device)
        for t in range(2, timesteps):
            active_mask = active_mask * (1 - dones[:, t])
            if float(active_mask.sum()) == 0:
                break
            predicted = torch.sigmoid(decoder_forward(nets["decoder"], z))
            diffs = active_mask * ((states[:, t] - predicted) ** 2).mean(dim=-1).mean(dim=-1).mean(dim=-1)
            mse.append(float(torch.mean(diffs) * bsz / torch.sum(active_mask)))
            mse_std.append(float(torch.std(diffs) * bsz / torch.sum(active_mask)))
            r_diffs = active_mask * (rewards[:, t].sum(-1) - reward_forward(nets["reward_predictor"], z).sum(-1)) ** 2
            rew.append(float(torch.mean(r_diffs) * bsz / torch.sum(active_mask)))
            rew_std.append(float(torch.std(r_diffs) * bsz / torch.sum(active_mask)))
            z = transition_forward(nets["transition"], z, eye[actions[:, t]], training=False)
    return mse, mse_std, rew, rew_std


def compute_rollout_reward(nets, z, num_actions, *, training=False, lookahead=2, rollout_depth=12,
                           negative_positive_tradeoff=10.0, uniform_source=None):
    """reference main.py:455-489 with rollout_policy='noop': best plan score of the A^2 beam started at z [1,L,H,W].
    uniform_source(shape) supplies the Bernoulli uniforms in train mode (None: torch.rand)."""
    assert lookahead == 2
    width = num_actions ** lookahead
    eye = torch.eye(num_actions, dtype=z.dtype, device=z.device)
    plans = torch.as_tensor([[i, j] + [0] * (rollout_depth - lookahead)
                             for i in range(num_actions) for j in range(num_actions)], device=z.device)
    with torch.no_grad():
        z = z.repeat(width, 1, 1, 1)
        cumulative = reward_forward(nets["reward_predictor"], z).clone()
        for t in range(rollout_depth):
            u = uniform_source(tuple(z.shape)) if (training and uniform_source is not None) else None
            z = transition_forward(nets["transition"], z, eye[plans[:, t]], training=training, uniforms=u)
            cumulative += reward_forward(nets["reward_predictor"], z)
        cumulative[:, 0] *= negative_positive_tradeoff
        return cumulative.sum(dim=1).max(dim=0)[0]


def choose_action(nets, z, num_actions, *, training=False, rollout_depth=12, uniform_source=None):
    """One decision of play(), reference main.py:356-368: (argmax action, per-action scores [A])."""
    eye = torch.eye(num_actions, dtype=z.dtype, device=z.device)
    scores = []
    with torch.no_grad():
        for a in range(num_actions):
            u = uniform_source(tuple(z.shape)) if (training and uniform_source is not None) else None
            z_a = transition_forward(nets["transition"], z, eye[a:a + 1], training=training, uniforms=u)
            scores.append(compute_rollout_reward(nets, z_a, num_actions, training=training,
                                                 rollout_depth=rollout_depth, uniform_source=uniform_source))
    scores = torch.stack(scores)
    return int(torch.argmax(scores)), scores


# --------------------------------------------------------------------------------------------------------------
# parameter initialisation in the reference's construction order (so seeded weights match bit for bit)
# --------------------------------------------------------------------------------------------------------------
def _conv_init(cout, cin, k=3, transposed=False):
    """nn.Conv2d / nn.ConvTranspose2d default init (kaiming_uniform(a=sqrt(5)) + uniform bias), same RNG order."""
    shape = (cin, cout, k, k) if transposed else (cout, cin, k, k)
    w = torch.empty(shape)
    torch.nn.init.kaiming_uniform_(w, a=math.sqrt(5))
    fan_in = shape[1] * k * k
    bound = 1 / math.sqrt(fan_in)
    b = torch.empty(cout).uniform_(-bound, bound)
    return w, b


def _sn_init(sd, name, cout, cin):
    """SpectralNorm(nn.Conv2d(...)): conv init, then u, v ~ N(0,1) normalised (spectral_normalization.py:47-63)."""
    w, b = _conv_init(cout, cin)
    u = l2normalize(torch.empty(cout).normal_(0, 1))
    v = l2normalize(torch.empty(cin * 9).normal_(0, 1))
    sd[name + ".module.bias"] = b
    sd[name + ".module.weight_u"] = u
    sd[name + ".module.weight_v"] = v
    sd[name + ".module.weight_bar"] = w


def init_transition(latent, num_actions):
    """reference models.py:44-57"""
    sd = {}
    _sn_init(sd, "conv1", 128, latent + num_actions)
    _sn_init(sd, "conv2", 128, 128)
    _sn_init(sd, "conv3", 128, 128)
    _sn_init(sd, "conv4", 128, 128)
    _sn_init(sd, "conv5", 128, 256)
    sd["conv6.weight"], sd["conv6.bias"] = _conv_init(latent, 256)
    return sd


def init_encoder(latent, color_channels):
    """reference models.py:124-137 (bn_conv1 is created between conv1 and conv2 and consumes no RNG)."""
    sd = {}
    _sn_init(sd, "conv1", 128, color_channels * ENCODER_INPUT_FRAMES)
    sd["bn_conv1.weight"] = torch.ones(128)
    sd["bn_conv1.bias"] = torch.zeros(128)
    sd["bn_conv1.running_mean"] = torch.zeros(128)
    sd["bn_conv1.running_var"] = torch.ones(128)
    sd["bn_conv1.num_batches_tracked"] = torch.tensor(0)
    _sn_init(sd, "conv2", 128, 128)
    _sn_init(sd, "conv3", 128, 128)
    sd["conv4.weight"], sd["conv4.bias"] = _conv_init(latent, 128)
    return sd


def init_decoder(latent, color_channels):
    """reference models.py:254-268"""
    sd = {}
    sd["conv1.weight"], sd["conv1.bias"] = _conv_init(latent * 4, latent, transposed=True)
    sd["conv2.weight"], sd["conv2.bias"] = _conv_init(latent * color_channels, latent * 4, transposed=True)
    return sd


def init_reward_predictor(latent, num_rewards):
    """reference models.py:228-233"""
    sd = {}
    sd["conv1.weight"], sd["conv1.bias"] = _conv_init(32, latent)
    sd["conv2.weight"], sd["conv2.bias"] = _conv_init(num_rewards * 3, 32)
    return sd


def synthetic_batch(batch, horizon, channels, height, width, num_actions, num_rewards, seed=1234, p_done=0.0):
    """Synthetic trajectories of SURVEY.md section 8d (sparse binary frames, 5% non-zero rewards)."""
    g = torch.Generator().manual_seed(seed)
    states = (torch.rand(batch, horizon, channels, height, width, generator=g) < 0.15).float()
    r = torch.rand(batch, horizon, num_rewards, generator=g)
    rewards = torch.where(r < 0.025, -1.0, torch.where(r > 0.975, 1.0, 0.0))
    dones = (torch.rand(batch, horizon, generator=g) < p_done).float()
    actions = torch.randint(num_actions, (batch, horizon), generator=g)
    return states, rewards, dones, actions.numpy()


# --------------------------------------------------------------------------------------------------------------
# envs/minipacman.py:122-164 (same sampler in the other environments): replay-buffer clip sampling
# --------------------------------------------------------------------------------------------------------------
def get_trajectories_from_uniforms(episodes, uniforms, batch_size, timesteps, random_start=True):
    """reference envs/minipacman.py:137-163 with the two random draws of every clip taken from `uniforms`:
    uniforms[b, k] = (u0, u1), u0 -> random.choice(replay_buffer) (line 144), u1 -> np.random.randint(0, len - 3)
    (line 146).  A uniform u selects index min(int(float32(u) * float32(n)), n - 1) of n choices.
    episodes: list of (states [n,...], rewards [n,R], actions [n]).
    Returns (states, rewards, dones, actions, plan) with plan[b] = [(episode, start, duration), ...]."""
    f32 = np.float32

    def pick(u, n):
        return min(int(f32(u) * f32(n)), n - 1)
    states_b, rewards_b, dones_b, actions_b, plans = [], [], [], [], []
    for b in range(batch_size):
        states, rewards, actions, dones, plan = [], [], [], [], []
        remaining, k = timesteps, 0
        while remaining > 0:
            e = pick(uniforms[b][k][0], len(episodes))
            sel_s, sel_r, sel_a = episodes[e]
            start = pick(uniforms[b][k][1], len(sel_s) - 3) if random_start else 0
            end = min(start + remaining, len(sel_s) - 1)
            duration = end - start
            states.extend(sel_s[start:end])
            rewards.extend(sel_r[start:end])
            actions.extend(sel_a[start:end])
            dones.extend([False for _ in range(duration - 1)] + [True])
            remaining -= duration
            plan.append((e, start, duration))
            k += 1
        states_b.append(np.array(states)); rewards_b.append(np.array(rewards))
        dones_b.append(np.array(dones)); actions_b.append(np.array(actions)); plans.append(plan)
    return np.array(states_b), np.array(rewards_b), np.array(dones_b), np.array(actions_b), plans
