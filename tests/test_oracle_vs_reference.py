"""CPU tests that need the reference checkout (/root/reference; skipped on the GPU box, where it does not exist):
the oracle restatement against the LIVE reference classes, and checkpoint compatibility of the drop-in modules."""
import copy

import pytest
import torch

from oracle import restated as R
from oracle import shims

pytestmark = pytest.mark.skipif(not shims.reference_available(), reason="reference checkout not present")


@pytest.fixture(scope="module")
def ref():
    return shims.load_reference_modules()


def _build(ref, C=3, A=5, Rw=2, seed=3):
    torch.manual_seed(seed)
    with shims.cpu_cuda_noop():
        m = ref["models"]
        nets = {"encoder": m.Encoder(16, C), "decoder": m.Decoder(16, C), "reward_predictor": m.RewardPredictor(16, Rw),
                "transition": m.Transition(16, A)}
    shims.apply_legacy_circular(nets["transition"])
    return nets


def test_restatement_matches_live_reference_forward_and_sn_state(ref):
    nets = _build(ref)
    sds = {k: copy.deepcopy(v.state_dict()) for k, v in nets.items()}
    g = torch.Generator().manual_seed(0)
    x = (torch.rand(2, 3, 3, 9, 11, generator=g) < 0.2).float()
    z = torch.rand(2, 16, 9, 11, generator=g)
    a = torch.eye(5)[torch.randint(5, (2,), generator=g)]
    u = torch.rand(2, 16, 9, 11, generator=g)
    with torch.no_grad(), shims.cpu_cuda_noop():
        for _ in range(2):  # two calls: the power iteration state must advance identically
            ze_ref = nets["encoder"](x)
            ze = R.encoder_forward(sds["encoder"], x)
            torch.testing.assert_close(ze, ze_ref, rtol=1e-5, atol=1e-6)
            with shims.injected_bernoulli(lambda shape: u):
                zt_ref = nets["transition"](z, a)
            zt = R.transition_forward(sds["transition"], z, a, training=True, uniforms=u)
            assert (zt != zt_ref).float().mean().item() < 1e-3
        torch.testing.assert_close(R.decoder_forward(sds["decoder"], z), nets["decoder"](z), rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(R.reward_forward(sds["reward_predictor"], z), nets["reward_predictor"](z), rtol=1e-5,
                                   atol=1e-5)
    for k, v in nets["transition"].state_dict().items():
        torch.testing.assert_close(sds["transition"][k], v, rtol=1e-5, atol=1e-7)


def test_checkpoints_load_both_ways(ref, tmp_path):
    """model-*.pth written by the reference (main.py:137-141) load into the drop-in modules and vice versa."""
    from scm_gan_b200.train_step import build_nets
    rnets = _build(ref, seed=5)
    ours = build_nets(3, 5, 2, seed=0)
    for name, rn in rnets.items():
        path = tmp_path / f"model-{name}.pth"
        torch.save(rn.state_dict(), path)
        sd = torch.load(path)
        missing = ours[name].load_state_dict(sd, strict=True)
        assert not missing.missing_keys and not missing.unexpected_keys
        for k, v in ours[name].state_dict().items():
            assert torch.equal(v.cpu(), sd[k]), f"{name}.{k}"
        back = rn.load_state_dict({k: v.cpu() for k, v in ours[name].state_dict().items()}, strict=True)
        assert not back.missing_keys and not back.unexpected_keys
    disc_ref = ref["models"].Discriminator.__new__(ref["models"].Discriminator)
    with shims.cpu_cuda_noop():
        disc_ref.__init__()
    from scm_gan_b200.train_step import import_dropin_models
    disc = import_dropin_models().Discriminator()
    assert list(disc.state_dict().keys()) == list(disc_ref.state_dict().keys())
