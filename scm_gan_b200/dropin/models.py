"""World-model networks with the reference's module interface (reference models.py), executed by hand-written
sm_100a kernels through the scmgan::* custom ops.

Put this directory first on sys.path and the reference's `main.py` runs unchanged (`import models`):
  Encoder(latent_size, color_channels)(x[B,3,C,H,W])              -> z [B,L,H,W]            models.py:123-157
  Transition(latent_size, num_actions)(z, onehot, return_all=...) -> z' in {0,1} (train)    models.py:43-119
  Decoder(latent_size, color_channels)(z, visualize=False)        -> logits [B,C,H,W]       models.py:253-291
  RewardPredictor(latent_dim, num_rewards)(z, visualize=False)    -> [B,R]                  models.py:226-250
Parameter names, shapes, construction order (RNG consumption) and state_dict keys are identical to the reference,
so checkpoints (`model-*.pth`, main.py:79-90,133-141) load both ways.  The nn.Conv2d children only *hold* the
parameters; the arithmetic of Encoder/Transition/Decoder happens in scm_gan_b200/csrc.

Interface-only classes of the reference (Discriminator, Inverter, RGBDecoder, GaussianSmoothing, random_eps, norm)
are provided for import compatibility; main.py constructs Discriminator but never calls it.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from scm_gan_b200 import ops as _ops
from spectral_normalization import SpectralNorm
from coordconv import CoordConv2d  # noqa: F401  (reference models.py:15 imports it)
from spatial_recurrent import CSRN  # noqa: F401  (reference models.py:14 imports it)

NOISE_DIM = 3
ENCODER_INPUT_FRAMES = 3


def _to_device(module):
    """The reference's constructors end in self.cuda() (models.py:57,137,233,268)."""
    if torch.cuda.is_available():
        module.cuda()
    return module


def random_eps(p=0.5, batch_size=32, height=64, width=64, channels=NOISE_DIM):
    shape = (batch_size, height, width, channels)
    return torch.bernoulli(torch.full(shape, float(p))).cuda()


class DifferentiableBernoulliSampler(torch.autograd.Function):
    """Sample in forward, identity in backward (reference models.py:30-40).  Transition fuses this into the
    conv6 epilogue; the class is kept for code that imports it."""

    @staticmethod
    def forward(ctx, x):
        return torch.bernoulli(x)

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output


def _sn_params(layers):
    mods = [l.module for l in layers]
    return ([m.weight_bar for m in mods], [m.bias for m in mods], [m.weight_u for m in mods],
            [m.weight_v for m in mods])


class Transition(nn.Module):
    def __init__(self, latent_size, num_actions):
        super().__init__()
        self.latent_size = latent_size
        hid = 128

        def circ(cin, cout):  # wrap-by-one 3x3 conv: the semantics the reference's `padding=2, 'circular'` had
            return nn.Conv2d(cin, cout, (3, 3), stride=1, padding=1, padding_mode='circular')

        self.conv1 = SpectralNorm(circ(latent_size + num_actions, hid))
        self.conv2 = SpectralNorm(circ(hid, hid))
        self.conv3 = SpectralNorm(circ(hid, hid))
        self.conv4 = SpectralNorm(circ(hid, hid))
        self.conv5 = SpectralNorm(circ(2 * hid, hid))   # input: cat[act4, skip2]
        self.conv6 = circ(2 * hid, latent_size)         # input: cat[act5, skip1]; not spectrally normalised
        self._uniforms = None  # test hook: uniforms consumed by the next training-mode forward (U < p)
        # Philox {seed, offset} for the in-kernel Bernoulli draw (not part of the state_dict, like torch's generator)
        self.register_buffer("_rng_state", torch.tensor([torch.initial_seed() & 0x7FFFFFFFFFFFFFFF, 0],
                                                        dtype=torch.int64), persistent=False)
        _to_device(self)

    def power_iterations(self, calls):
        """Run the power iterations of the next `calls` forward calls ahead of time, in one kernel launch: returns
        sigma [calls, 5] (row i = what call i would compute, spectral_normalization.py:28-34) and leaves u, v advanced
        `calls` times, exactly as `calls` forward calls would.  The iteration reads nothing but the weights, so this
        is only valid while they do not change (within one training iteration).  Pass row i - or several rows, when
        the batch folds the samples of several calls - as `sigma=` to forward()."""
        wbar, _, u, v = _sn_params([self.conv1, self.conv2, self.conv3, self.conv4, self.conv5])
        sigma = torch.empty((calls, len(wbar)), dtype=torch.float32, device=wbar[0].device)
        with torch.no_grad():
            _ops.K.spectral_norm_fwd_n([w.detach() for w in wbar], [t.detach() for t in u], [t.detach() for t in v],
                                       sigma)
        return sigma

    def forward(self, s, a, return_all=False, sigma=None):
        """sigma (optional, not part of the reference signature): rows of power_iterations().  One row [5]: this call
        uses it instead of running its own power iteration.  Several rows [S, 5]: the batch consists of S equal
        segments belonging to S different calls of the reference (the main rollout step and the counterfactual
        rollouts of main.py:242-283 folded into one batch); segment s is normalised with row s."""
        assert s.shape[0] == a.shape[0]
        wbar, bias, u, v = _sn_params([self.conv1, self.conv2, self.conv3, self.conv4, self.conv5])
        if sigma is None:
            with torch.no_grad():
                sigma = torch.ops.scmgan.spectral_norm_update(wbar, u, v)
        uniforms = None
        if self.training:
            uniforms, self._uniforms = self._uniforms, None
            if uniforms is None:
                _ops._RNG_SOURCE.append(self._rng_state)  # sample inside the conv6 epilogue (Philox4x32-10)
        _ops._UV_SOURCE.append((u, v))
        try:
            out = torch.ops.scmgan.transition_fwd(s, a, wbar, bias, sigma, self.conv6.weight, self.conv6.bias,
                                                  uniforms, self.training)
        finally:
            _ops._UV_SOURCE.clear()
            _ops._RNG_SOURCE.clear()
        x = out[0]
        if not self.training:
            x = x.detach()  # thresholding carries no gradient (reference models.py:112)
        if return_all:
            zin, buf6, buf5, act3 = out[2:6]
            hid = act3.shape[3]

            def nchw(plane, c0):
                return plane[:, 1:-1, 1:-1, c0:c0 + hid].permute(0, 3, 1, 2).float()
            return (nchw(buf6, hid), nchw(buf5, hid), nchw(act3, 0), nchw(buf5, 0), nchw(buf6, 0), x)
        return x


class Encoder(nn.Module):
    def __init__(self, latent_size, color_channels):
        super().__init__()
        self.latent_size = latent_size
        self.color_channels = color_channels
        hid = 128
        self.conv1 = SpectralNorm(nn.Conv2d(color_channels * ENCODER_INPUT_FRAMES, hid, (3, 3), stride=1, padding=1))
        self.bn_conv1 = nn.BatchNorm2d(hid)  # registered but never applied, as in the reference (models.py:130)
        self.conv2 = SpectralNorm(nn.Conv2d(hid, hid, (3, 3), stride=1, padding=1))
        self.conv3 = SpectralNorm(nn.Conv2d(hid, hid, (3, 3), stride=1, padding=1))
        self.conv4 = nn.Conv2d(hid, latent_size, (3, 3), stride=1, padding=1)
        _to_device(self)

    def forward(self, x):
        batch_size, frames, channels, height, width = x.shape
        x = x.view(batch_size, frames * channels, height, width)
        wbar, bias, u, v = _sn_params([self.conv1, self.conv2, self.conv3])
        with torch.no_grad():
            sigma = torch.ops.scmgan.spectral_norm_update(wbar, u, v)
        _ops._UV_SOURCE.append((u, v))
        try:
            out = torch.ops.scmgan.encoder_fwd(x, wbar, bias, sigma, self.conv4.weight, self.conv4.bias)
        finally:
            _ops._UV_SOURCE.clear()
        return out[0]


class Decoder(nn.Module):
    def __init__(self, latent_size, color_channels):
        super().__init__()
        self.latent_size = latent_size
        self.color_channels = color_channels
        self.conv1 = nn.ConvTranspose2d(latent_size, latent_size * 4, (3, 3), stride=1, padding=1)
        self.conv2 = nn.ConvTranspose2d(latent_size * 4, latent_size * color_channels, (3, 3), stride=1, padding=1)
        self._fold = None  # (key, folded weight, folded bias) shared by the decoder calls of one iteration
        _to_device(self)

    def _folded_last_layer(self):
        """conv2 with the sum over latent groups folded into its weights.  Every unrolled step of an iteration decodes
        with the same weights, so the fold (and its backward) is done once per autograd graph: the cached tensors are
        dropped when their gradient arrives, or when the parameters change version (load_state_dict, optimizer)."""
        w2, b2 = self.conv2.weight, self.conv2.bias
        key = (w2._version, b2._version, w2.data_ptr(), torch.is_grad_enabled())
        if self._fold is not None and self._fold[0] == key:
            return self._fold[1], self._fold[2]
        hid = w2.shape[0]
        w2f = w2.view(hid, self.latent_size, self.color_channels, 3, 3).sum(1)
        b2f = b2.view(self.latent_size, self.color_channels).sum(0)
        if w2f.requires_grad:
            w2f.register_hook(self._drop_fold)
            self._fold = (key, w2f, b2f)
        return w2f, b2f

    def _drop_fold(self, grad):
        self._fold = None
        return grad

    def __getstate__(self):  # the cache holds non-leaf tensors: never copied or pickled with the module
        state = self.__dict__.copy()
        state["_fold"] = None
        return state

    def forward(self, z_map, visualize=False):
        batch_size, latent_size, height, width = z_map.shape
        w2, b2 = self.conv2.weight, self.conv2.bias
        if visualize:
            # per-factor maps requested: run the un-folded second layer (latent_size*C output channels)
            x = torch.ops.scmgan.decoder_fwd(z_map, self.conv1.weight, self.conv1.bias, w2, b2)[0]
            x = x.view(batch_size, latent_size, self.color_channels, height, width)
            return torch.sum(x, dim=1), x[0]
        # sum over the latent groups commutes with the (linear) last layer: fold it into the weights, exactly
        # (up to fp reassociation).  Autograd broadcasts the folded gradient back to all groups.
        w2f, b2f = self._folded_last_layer()
        return torch.ops.scmgan.decoder_fwd(z_map, self.conv1.weight, self.conv1.bias, w2f, b2f)[0]


    def pixel_loss_seq(self, z_map, target_bt, mask_bt):
        """Not part of the reference interface (used by scm_gan_b200.train_step): decode the latents of all T rollout
        steps (z_map [T*B, L, H, W], t-major) and return the T reconstruction terms of main.py:188-197,
        mean_b(mask[b,t] * mean_chw BCE(sigmoid(decoder(z)), target[b,t])), with sigmoid + BCE + the means evaluated in
        the epilogue of the last convolution (scmgan_decoder_bce_fwd).  target_bt [B, T, C, H, W], mask_bt [B, T]."""
        w2f, b2f = self._folded_last_layer()
        return torch.ops.scmgan.decoder_bce_seq(z_map, self.conv1.weight, self.conv1.bias, w2f, b2f, target_bt,
                                                mask_bt)[0]


class RewardPredictor(nn.Module):
    # Each reward is a per-pixel 3-way classification (+1 / 0 / -1) summed over the map (reference models.py:226-250)
    def __init__(self, latent_dim, num_rewards):
        super().__init__()
        self.conv1 = nn.Conv2d(latent_dim, 32, (3, 3), stride=1, padding=0)
        self.conv2 = nn.Conv2d(32, num_rewards * 3, (3, 3), stride=2, padding=0)
        _to_device(self)

    def forward(self, x, visualize=False):
        out = torch.ops.scmgan.reward_fwd(x, self.conv1.weight, self.conv1.bias, self.conv2.weight, self.conv2.bias)
        if visualize:
            return out[0], out[1]
        return out[0]


class Inverter(nn.Module):
    """Interface-only (reference models.py:167-190; its forward references an undefined variable)."""

    def __init__(self, latent_size):
        super().__init__()
        self.latent_size = latent_size
        self.conv1 = nn.Conv2d(latent_size * 2, 32, (3, 3), stride=1, padding=1)
        self.conv2 = SpectralNorm(nn.Conv2d(32, NOISE_DIM, (3, 3), stride=1, padding=0))
        _to_device(self)

    def forward(self, s_curr, s_next, a):
        x = torch.cat([s_curr, s_next], dim=1)
        x = F.leaky_relu(self.conv1(x))
        return torch.sigmoid(self.conv2(x))


class Discriminator(nn.Module):
    """Constructed and checkpointed by main.py (76,140) but never called; parameters match models.py:195-209."""

    def __init__(self):
        super().__init__()
        self.conv1 = SpectralNorm(nn.Conv2d(NOISE_DIM, 32, (3, 3), stride=2, padding=0))
        self.conv2 = SpectralNorm(nn.Conv2d(32, 32, (3, 3), stride=2, padding=0))
        self.conv3 = nn.Conv2d(32, 32, (3, 3), stride=2, padding=0)
        self.fc1 = nn.Linear(32 * 7 * 7, 1)
        _to_device(self)

    def forward(self, x):
        x = F.leaky_relu(self.conv1(x))
        x = F.leaky_relu(self.conv2(x))
        x = F.leaky_relu(self.conv3(x))
        return F.leaky_relu(self.fc1(x.flatten(1)))


class RGBDecoder(nn.Module):
    """Identity with a registered background parameter (reference models.py:294-310)."""

    def __init__(self, color_channels=3, img_size=256):
        super().__init__()
        bg = torch.zeros((color_channels, img_size, img_size))
        self.bg = nn.Parameter(bg.cuda() if torch.cuda.is_available() else bg)

    def forward(self, x, enable_bg=True):
        return x


class GaussianSmoothing(nn.Module):
    """Depthwise Gaussian filter (reference models.py:315-378); not used by main.py."""

    def __init__(self, channels, kernel_size, sigma, dim=2):
        super().__init__()
        if dim not in (1, 2, 3):
            raise RuntimeError('Only 1, 2 and 3 dimensions are supported. Received {}.'.format(dim))
        self.padding = [int(kernel_size / 2)] * dim
        sizes, sigmas = [kernel_size] * dim, [sigma] * dim
        grids = torch.meshgrid([torch.arange(s, dtype=torch.float32) for s in sizes], indexing="ij")
        kernel = torch.ones(())
        for size, std, grid in zip(sizes, sigmas, grids):
            mean = (size - 1) / 2
            kernel = kernel * (1 / (std * math.sqrt(2 * math.pi)) * torch.exp(-((grid - mean) / (2 * std)) ** 2))
        kernel = kernel / kernel.sum()
        kernel = kernel.view(1, 1, *kernel.shape).repeat(channels, *[1] * (kernel.dim() + 1))
        self.register_buffer('weight', kernel)
        self.groups = channels
        self.conv = {1: F.conv1d, 2: F.conv2d, 3: F.conv3d}[dim]
        _to_device(self)

    def forward(self, input):
        return self.conv(input, weight=self.weight, groups=self.groups, padding=self.padding)


def norm(x):
    """Normalise a batch of latent points to the unit hypersphere (reference models.py:382-385)."""
    n = torch.norm(x, p=2, dim=1)
    return x / (n.expand(1, -1).t() + .0001)
