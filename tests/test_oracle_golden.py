"""CPU tests: the oracle restatement (oracle/restated.py) is pinned against golden vectors produced by the
UNMODIFIED reference modules (oracle/make_golden.py, run in the build container)."""
import copy
import os

import pytest
import torch

from oracle import restated as R

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
CONFIGS = ["minipacman", "pong64", "sc2"]


def load(name):
    return torch.load(os.path.join(GOLDEN, f"{name}.pt"), weights_only=False)


def check_summary(t, s, rtol, what):
    t = t.detach().float().flatten()
    assert t.numel() == s["numel"], what
    ref = s["val"]
    got = t[s["idx"]]
    scale = max(s["norm"] / max(s["numel"], 1) ** 0.5, 1e-12)
    err = (got - ref).abs().max().item()
    assert err <= rtol * scale + 1e-9, f"{what}: sample err {err:.3e} vs rms {scale:.3e}"
    assert abs(t.norm().item() - s["norm"]) <= rtol * max(s["norm"], 1e-9), f"{what}: norm"


def fresh_nets(cfg, requires_grad=False):
    torch.manual_seed(cfg["seed"])
    nets = {"encoder": R.init_encoder(16, cfg["C"]), "decoder": R.init_decoder(16, cfg["C"]),
            "reward_predictor": R.init_reward_predictor(16, cfg["R"]), "transition": R.init_transition(16, cfg["A"])}
    if requires_grad:
        for sd in nets.values():
            for k, v in sd.items():
                if v.dtype.is_floating_point and not (k.endswith("_u") or k.endswith("_v") or "bn_conv1" in k):
                    v.requires_grad_(True)
    return nets


@pytest.mark.parametrize("name", CONFIGS)
def test_seeded_init_matches_reference_constructors(name):
    g = load(name)
    nets = fresh_nets(g["config"])
    for net, summ in g["weights"].items():
        for k, s in summ.items():
            check_summary(nets[net][k], s, 0.0, f"{net}.{k}")


@pytest.mark.parametrize("name", CONFIGS)
def test_module_forward_matches_reference(name):
    g = load(name)
    cfg, m, inp = g["config"], g["modules"], g["inputs"]
    nets = fresh_nets(cfg)
    with torch.no_grad():
        z = R.encoder_forward(nets["encoder"], inp["states"][:, 0:3])
        torch.testing.assert_close(z, m["encoder_z"], rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(R.decoder_forward(nets["decoder"], m["zreal"]), m["decoder_logits"], rtol=1e-5,
                                   atol=1e-5)
        torch.testing.assert_close(R.decoder_forward(nets["decoder"], m["zreal"], visualize=True)[1],
                                   m["decoder_logits_vis"], rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(R.reward_forward(nets["reward_predictor"], m["zreal"]), m["reward"], rtol=1e-5,
                                   atol=1e-5)
        outs = R.transition_forward(nets["transition"], m["zin"], m["onehot"], training=True, uniforms=m["u0"],
                                    return_all=True)
        for got, ref in zip(outs[:5], m["transition_all"][:5]):
            torch.testing.assert_close(got, ref, rtol=1e-5, atol=1e-5)
        assert (outs[5] != m["transition_all"][5]).float().mean().item() < 1e-3  # sampled bits (u ~ p ties only)
        ze = R.transition_forward(nets["transition"], m["zreal"], m["onehot"], training=False)
        assert (ze != m["transition_eval"]).float().mean().item() < 1e-3
    # spectral-norm state advanced exactly like the reference's .data updates
    for net, st in g["sn_state"].items():
        for k, v in st.items():
            torch.testing.assert_close(nets[net][k], v, rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("name", CONFIGS)
def test_training_step_loss_and_grads_match_reference(name):
    g = load(name)
    cfg, inp, st = g["config"], g["inputs"], g["step"]
    nets = fresh_nets(cfg, requires_grad=True)
    loss, terms, zfin = R.train_step_loss(
        nets, inp["states"], inp["rewards"], inp["dones"], inp["actions"].numpy(), num_actions=cfg["A"],
        theta=st["theta"], uniforms=copy.deepcopy(st["uniforms"]), enable_disentanglement=True,
        enable_action_control=True, cf_now=True, counterfactual_horizon=st["cf_horizon"],
        cf_indices=st["cf_indices"], cf_perm=st["cf_perm"])
    assert abs(loss.item() - st["loss"].item()) <= 1e-5 * abs(st["loss"].item())
    for k, v in st["terms"].items():
        assert abs(terms[k].item() - v.item()) <= 1e-4 * abs(v.item()) + 1e-7, k
    assert (zfin != st["z_final"]).float().mean().item() < 1e-3
    loss.backward()
    for net, summ in st["grads"].items():
        for k, s in summ.items():
            grad = nets[net][k].grad
            assert grad is not None, f"{net}.{k} has no grad"
            check_summary(grad, s, 2e-3, f"grad {net}.{k}")


def test_layers_coordconv_csrn():
    g = torch.load(os.path.join(GOLDEN, "layers.pt"), weights_only=False)
    cc = g["coordconv"]
    y = R.coordconv_forward(cc["state"]["conv.weight"], cc["state"]["conv.bias"], cc["x"], padding=1)
    torch.testing.assert_close(y, cc["y"], rtol=1e-5, atol=1e-5)
    cs = g["csrn"]
    y = R.csrn_forward(cs["state"], cs["x"])
    torch.testing.assert_close(y, cs["y"], rtol=1e-4, atol=1e-3)


def test_mpc_planner_matches_reference_play_loop():
    """oracle.restated.choose_action / compute_rollout_reward against the unmodified reference's
    `compute_rollout_reward` driven as in `play()` (main.py:356-368, 455-489; oracle/make_golden_planner.py)."""
    g = load("planner")
    nets = fresh_nets(g["config"])
    for k, v in g["sn_before"].items():
        assert torch.equal(nets["transition"][k], v), k
    best, scores = R.choose_action(nets, g["z0"], g["config"]["A"], training=False)
    ref = g["scores"]
    assert best == g["best_action"]
    assert torch.allclose(scores, ref, rtol=1e-4, atol=1e-3), (scores, ref)
    for k, v in g["sn_after"].items():  # 65 Transition calls advanced the power iteration identically
        assert torch.allclose(nets["transition"][k], v, rtol=1e-4, atol=1e-5), k
