"""TEST INFRASTRUCTURE: runs the reference's own, UNMODIFIED `main.py` (LilJing/scm-gan) on the drop-in modules.

`python main.py --env minipacman ...` needs, besides the four layer/model modules this repository replaces, a handful of
packages that cannot be installed offline (imutil, logutil, gym, gym_minipacman, matplotlib).  This harness supplies them
as stand-ins OUTSIDE the product, puts `scm_gan_b200/dropin` first on sys.path so that
`import models` / `spectral_normalization` / `coordconv` / `spatial_recurrent` resolve to the sm_100a-backed modules, and
then executes the reference file where it lies (`baseline/_ref/main.py`, a git-ignored verbatim copy of
/root/reference made by prepare_ref_copy(); the GPU box has no /root/reference):

  as a script   python tests/ref_main_harness.py --env minipacman --train-iters 30 ...     (runpy, run_name="__main__")
  as a module   run_main([...], iters_per_video=25)   loads main.py under another module name, optionally lowers its
                ITERS_PER_VIDEO constant (2000, main.py:53) and replaces evaluate() (video / plot output through
                imutil / matplotlib / pandas.plot, main.py:315-322) by a stub so that the checkpointing branch of the
                training loop (main.py:133-141) and --load-from / --start-iter (main.py:79-90) can be exercised in
                seconds.  train() / main() themselves run unchanged.

Shims (all documented in SURVEY.md section 8c): stub modules; a synthetic `gym_minipacman` environment producing
15x19x3 frames (scm_gan_b200.synthetic.MovingDotsEnv dynamics); torch<=1.4 behaviour of clip_grad_value_ on a network
without gradients (main.py:289 at horizon 3).
"""
import importlib.util
import os
import runpy
import shutil
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_COPY = os.path.join(ROOT, "baseline", "_ref")
DROPIN = os.path.join(ROOT, "scm_gan_b200", "dropin")


def prepare_ref_copy(src="/root/reference"):
    """Verbatim copy of the reference's Python files into the git-ignored baseline/_ref (travels to the GPU box)."""
    if os.path.isfile(os.path.join(REF_COPY, "main.py")):
        return True
    if not os.path.isfile(os.path.join(src, "main.py")):
        return False
    os.makedirs(REF_COPY, exist_ok=True)
    for name in os.listdir(src):
        p = os.path.join(src, name)
        if name.endswith(".py"):
            shutil.copy2(p, os.path.join(REF_COPY, name))
    shutil.copytree(os.path.join(src, "envs"), os.path.join(REF_COPY, "envs"), dirs_exist_ok=True)
    return True


class RecordingTimeSeries:
    """Stand-in for logutil.TimeSeries (main.py:130,184,196,297) that keeps what main.py collects."""
    instances = []

    def __init__(self, *a, **k):
        self.series = {}
        RecordingTimeSeries.instances.append(self)

    def collect(self, name, value):
        self.series.setdefault(name, []).append(float(value))   # the reference's logger reads the value too (D2H)

    def print_every(self, *a, **k):
        pass

    def __str__(self):
        return "TimeSeries(%s)" % ", ".join(f"{k}: {v[-1]:.4f}" for k, v in self.series.items() if v)


def _install_environment():
    import numpy as np
    from oracle import shims
    shims.install_stub_modules()
    sys.modules["logutil"].TimeSeries = RecordingTimeSeries
    from scm_gan_b200.synthetic import MovingDotsEnv

    class _Space:
        def __init__(self, n, rng):
            self.n, self.rng = n, rng

        def sample(self):
            return int(self.rng.randint(self.n))

    class MiniPacman:
        """Synthetic stand-in for gym_minipacman.envs.minipacman_env.MiniPacman: 15x19x3 HWC frames, 5 actions."""

        def __init__(self):
            self._env = MovingDotsEnv(3, 15, 19, 5, 2, seed=int(np.random.randint(1 << 30)), episode_length=10 ** 9)
            self.action_space = _Space(5, np.random.RandomState(int(np.random.randint(1 << 30))))

        def reset(self):
            return self._env.reset().transpose(1, 2, 0).copy()

        def step(self, action):
            frame, reward, done, info = self._env.step(action)
            return frame.transpose(1, 2, 0).copy(), reward, done, info

    class ALE:
        def __init__(self, *a, **k):
            pass

    mod = sys.modules["gym_minipacman.envs.minipacman_env"]
    mod.MiniPacman, mod.ALE = MiniPacman, ALE
    # shim 3: clip_grad_value_ on a network that received no gradient was a no-op on the author's torch
    import torch
    clip_mod = torch.nn.utils.clip_grad
    if not getattr(clip_mod.clip_grad_value_, "_scm_legacy", False):
        orig = clip_mod.clip_grad_value_

        def legacy(parameters, clip_value, *a, **k):
            params = [p for p in parameters if p.grad is not None]
            if params:
                return orig(params, clip_value, *a, **k)
        legacy._scm_legacy = True
        clip_mod.clip_grad_value_ = legacy
        torch.nn.utils.clip_grad_value_ = legacy
    if USE_REFERENCE_MODELS:
        _reference_model_shims()


def _reference_model_shims():
    """Control arm only: the reference's own models.py needs shim 2 (legacy circular pad-1 on every Transition) and,
    without a GPU, shim 4 (.cuda() no-op)."""
    import torch
    from oracle import shims
    if not torch.cuda.is_available():
        torch.nn.Module.cuda = lambda self, *a, **k: self
        torch.Tensor.cuda = lambda self, *a, **k: self
    import models
    if not getattr(models.Transition, "_scm_legacy_pad", False):
        orig_init = models.Transition.__init__

        def init(self, *a, **k):
            orig_init(self, *a, **k)
            shims.apply_legacy_circular(self)
        models.Transition.__init__ = init
        models.Transition._scm_legacy_pad = True


# SCMGAN_HARNESS_MODELS=reference: main.py imports the reference's OWN models.py / layer modules (stock torch ops, with
# the legacy-circular-padding shim) instead of the drop-ins - the control arm of the comparison, and the way to check the
# harness itself on a machine without a GPU.
USE_REFERENCE_MODELS = os.environ.get("SCMGAN_HARNESS_MODELS", "dropin") == "reference"


def _paths():
    order = (DROPIN, ROOT, REF_COPY) if USE_REFERENCE_MODELS else (REF_COPY, ROOT, DROPIN)
    for p in order:   # final order (default): dropin, repo root, reference copy
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    for name in ("models", "spectral_normalization", "coordconv", "spatial_recurrent", "datasource", "main"):
        sys.modules.pop(name, None)   # a previous import of the reference's own modules must not shadow the drop-ins


def run_main(argv, iters_per_video=None, stub_evaluate=True, spy=None):
    """Load baseline/_ref/main.py as a module (its argparse sees `argv`) and call its main().  Returns the module."""
    if not prepare_ref_copy():
        raise RuntimeError("reference copy not available")
    _paths()
    _install_environment()
    old_argv = sys.argv
    sys.argv = ["main.py"] + list(argv)
    try:
        spec = importlib.util.spec_from_file_location("scm_ref_main", os.path.join(REF_COPY, "main.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        import models
        want = REF_COPY if USE_REFERENCE_MODELS else DROPIN
        assert os.path.abspath(models.__file__).startswith(want), f"main.py imported {models.__file__}"
        if iters_per_video is not None:
            mod.ITERS_PER_VIDEO = iters_per_video
        if stub_evaluate:
            mod.evaluate = lambda *a, **k: print("[harness] evaluate() skipped (video / plot output)")
        if spy is not None:
            orig_train = mod.train

            def train(*a, **k):
                spy(*a, **k)
                return orig_train(*a, **k)
            mod.train = train
        mod.main()
        return mod
    finally:
        sys.argv = old_argv


def _dump_state_spy(path):
    """spy for run_main(): saves the state_dicts train() was entered with (after --load-from, main.py:79-90)."""
    def spy(latent_dim, datasource, num_actions, num_rewards, encoder, decoder, reward_predictor, discriminator,
            transition):
        import torch
        torch.save({"encoder": encoder.state_dict(), "decoder": decoder.state_dict(),
                    "transition": transition.state_dict(), "discriminator": discriminator.state_dict(),
                    "reward_predictor": reward_predictor.state_dict()}, path)
    return spy


def _report():
    import json
    ts = RecordingTimeSeries.instances[-1] if RecordingTimeSeries.instances else None
    if ts is not None:
        print("HARNESS_SERIES " + json.dumps({k: [v[0], v[-1], len(v)] for k, v in ts.series.items()}))
        print("HARNESS_FULL " + json.dumps(ts.series))
    print("HARNESS_MODELS " + os.path.abspath(sys.modules["models"].__file__))
    try:
        from scm_gan_b200 import kernels
        print("HARNESS_LAUNCHES %d" % kernels.launch_count())
    except Exception as e:  # control arm on a machine without the library
        print("HARNESS_LAUNCHES -1 (%s)" % type(e).__name__)


if __name__ == "__main__":
    # environment knobs (the command line belongs to main.py's own argparse):
    #   SCMGAN_HARNESS_MODELS=reference        control arm, see above
    #   SCMGAN_HARNESS_ITERS_PER_VIDEO=N       module mode with main.ITERS_PER_VIDEO lowered to N (checkpoint branch)
    #   SCMGAN_HARNESS_DUMP_STATE=file.pt      save the state_dicts train() starts from (checks --load-from)
    #   SCMGAN_HARNESS_SEED=S                  seed torch / numpy / random before main() (main.py itself never seeds)
    if not prepare_ref_copy():
        sys.exit("reference copy not available (baseline/_ref missing and /root/reference absent)")
    seed = os.environ.get("SCMGAN_HARNESS_SEED")
    if seed is not None:
        import random
        import numpy
        import torch
        torch.manual_seed(int(seed)); numpy.random.seed(int(seed)); random.seed(int(seed))
    ipv = os.environ.get("SCMGAN_HARNESS_ITERS_PER_VIDEO")
    dump = os.environ.get("SCMGAN_HARNESS_DUMP_STATE")
    if ipv is not None or dump is not None:
        run_main(sys.argv[1:], iters_per_video=int(ipv) if ipv else None,
                 spy=_dump_state_spy(dump) if dump else None)
    else:
        _paths()
        _install_environment()
        sys.argv = [os.path.join(REF_COPY, "main.py")] + sys.argv[1:]
        runpy.run_path(os.path.join(REF_COPY, "main.py"), run_name="__main__")
    _report()
