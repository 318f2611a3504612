"""TEST INFRASTRUCTURE ONLY - never imported by the product path (scm_gan_b200/).

Compatibility shims that let the *unmodified* reference (LilJing/scm-gan, /root/reference) be imported and
executed on this image (torch 2.11, no imutil/logutil/gym/matplotlib, possibly no GPU).  SURVEY.md section 8c
lists why each shim exists:

  1. sys.modules stubs for imutil, logutil, gym, gym_minipacman, matplotlib (reference models.py:11-12,
     main.py:13-22, datasource.py / envs/*.py imports) - none of them does model arithmetic;
  2. legacy circular padding: `padding=2, padding_mode='circular'` (reference models.py:51-56) expanded to one
     pixel per side on torch 1.1-1.4, the only semantics under which models.py:95 (`torch.cat` with the
     un-padded skip) ever ran;
  3. clip_grad_value_ on a network that received no gradient (reference main.py:287-290 at horizon 3) was a no-op;
  4. on a GPU-less host `.cuda()` is made a no-op so the reference constructors (models.py:57,137,233,268) run.
"""
import contextlib
import os
import sys
import types

import torch
import torch.nn as nn

REFERENCE_DIR = os.environ.get("SCMGAN_REFERENCE_DIR", "/root/reference")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_DIR, "models.py"))


class _TimeSeries:
    """Stand-in for logutil.TimeSeries (reference main.py:130,184,196,297): records values, prints nothing."""

    def __init__(self, *a, **k):
        self.series = {}

    def collect(self, name, value):
        self.series.setdefault(name, []).append(value)

    def print_every(self, *a, **k):
        pass

    def __str__(self):
        return "TimeSeries(%d series)" % len(self.series)


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__scm_stub__ = True
    return m


def install_stub_modules():
    """Idempotently register stub modules for the reference's un-installable imports."""
    def noop(*a, **k):
        return None

    class _Video:
        def __init__(self, *a, **k):
            pass

        def write_frame(self, *a, **k):
            pass

        def finish(self, *a, **k):
            pass

    stubs = {
        "imutil": _stub("imutil", show=noop, Video=_Video, VideoLoop=_Video, VideoMaker=_Video, get_pixels=noop,
                        encode_video=noop, load=noop),
        "logutil": _stub("logutil", TimeSeries=_TimeSeries, sparkline=lambda *a, **k: ""),
    }
    # gym & friends: only needed to import datasource/envs/main
    class _Env:
        pass

    class _Discrete:
        def __init__(self, n):
            self.n = n

    gym = _stub("gym", Env=_Env, make=noop)
    spaces = _stub("gym.spaces", Discrete=_Discrete, Box=_Discrete)
    discrete = _stub("gym.spaces.discrete", Discrete=_Discrete)
    gym.spaces = spaces
    spaces.discrete = discrete
    stubs.update({"gym": gym, "gym.spaces": spaces, "gym.spaces.discrete": discrete})
    gmp = _stub("gym_minipacman")
    gmp_envs = _stub("gym_minipacman.envs")
    gmp_env = _stub("gym_minipacman.envs.minipacman_env", MiniPacman=_Env, ALE=_Env)
    gmp.envs = gmp_envs
    gmp_envs.minipacman_env = gmp_env
    stubs.update({"gym_minipacman": gmp, "gym_minipacman.envs": gmp_envs,
                  "gym_minipacman.envs.minipacman_env": gmp_env})
    mpl = _stub("matplotlib", use=noop)
    plt = _stub("matplotlib.pyplot", figure=noop, plot=noop, savefig=noop, close=noop, clf=noop)
    mpl.pyplot = plt
    stubs.update({"matplotlib": mpl, "matplotlib.pyplot": plt})
    for name, mod in stubs.items():
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = mod


@contextlib.contextmanager
def cpu_cuda_noop():
    """Shim 4: make .cuda() a no-op while no CUDA device exists."""
    if torch.cuda.is_available():
        yield
        return
    mod_cuda, ten_cuda = nn.Module.cuda, torch.Tensor.cuda
    nn.Module.cuda = lambda self, *a, **k: self
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        nn.Module.cuda, torch.Tensor.cuda = mod_cuda, ten_cuda


_REF_MODULE_NAMES = ("models", "spectral_normalization", "coordconv", "spatial_recurrent")


def load_reference_modules():
    """Import the reference's layer library + models, unmodified, from REFERENCE_DIR.

    Returns a dict name -> module.  The modules are removed from sys.modules again so they never shadow the
    product's drop-in modules of the same names.
    """
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_DIR}")
    install_stub_modules()
    saved = {n: sys.modules.pop(n) for n in _REF_MODULE_NAMES if n in sys.modules}
    sys.path.insert(0, REFERENCE_DIR)
    try:
        with cpu_cuda_noop():
            mods = {n: __import__(n) for n in _REF_MODULE_NAMES}
    finally:
        sys.path.remove(REFERENCE_DIR)
        for n in _REF_MODULE_NAMES:
            sys.modules.pop(n, None)
        sys.modules.update(saved)
    return mods


def apply_legacy_circular(transition):
    """Shim 2: make every circular conv of a reference Transition wrap by one pixel per side."""
    for name in ("conv1", "conv2", "conv3", "conv4", "conv5", "conv6"):
        conv = getattr(transition, name)
        conv = getattr(conv, "module", conv)
        conv.padding = (1, 1)
        conv._reversed_padding_repeated_twice = (1, 1, 1, 1)
    return transition


def clip_grad_value_legacy(parameters, clip_value):
    """Shim 3: clip_grad_value_ that tolerates parameters without gradients (torch<=1.4 behaviour)."""
    params = [p for p in parameters if p.grad is not None]
    if params:
        torch.nn.utils.clip_grad_value_(params, clip_value)


@contextlib.contextmanager
def injected_bernoulli(uniform_source):
    """Shim 5: replace torch.bernoulli(p) by (U < p) with U drawn from `uniform_source(shape)` so that the
    reference and the implementation under test consume the same uniforms."""
    orig = torch.bernoulli

    def fake(p, *a, **k):
        u = uniform_source(tuple(p.shape)).to(p.device)
        return (u < p).to(p.dtype)

    torch.bernoulli = fake
    try:
        yield
    finally:
        torch.bernoulli = orig
