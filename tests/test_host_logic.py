"""CPU tests of the host side: drop-in module surface (names, state_dict keys, seeded init identical to the
reference), no-CPU-fallback behaviour of the custom ops, and the data-parallel gradient exchange on gloo (2 ranks)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def nets():
    from scm_gan_b200.train_step import build_nets
    return build_nets(3, 5, 2, seed=0)


def test_dropin_state_dict_and_seeded_init_match_reference(nets):
    g = torch.load(os.path.join(GOLDEN, "minipacman.pt"), weights_only=False)
    for net, summ in g["weights"].items():
        sd = nets[net].state_dict()
        assert {k for k, v in sd.items() if v.dtype.is_floating_point} == set(summ), net
        for k, s in summ.items():
            t = sd[k].detach().float().flatten().cpu()
            assert t.numel() == s["numel"] and torch.equal(t[s["idx"]], s["val"]), f"{net}.{k}"
    # parameter census of SURVEY.md a12 (C=3, A=5, R=2)
    count = {k: sum(p.numel() for p in m.parameters() if p.requires_grad) for k, m in nets.items()}
    assert count == {"encoder": 324368, "decoder": 36976, "reward_predictor": 6374, "transition": 798992}
    # u, v are parameters without gradient, exactly as in the reference (they appear in parameters()/state_dict())
    u = nets["transition"].conv1.module.weight_u
    assert isinstance(u, torch.nn.Parameter) and not u.requires_grad


def test_module_interface_names():
    from scm_gan_b200.train_step import import_dropin_models
    m = import_dropin_models()
    for name in ("Encoder", "Transition", "Decoder", "RewardPredictor", "Discriminator", "Inverter", "RGBDecoder",
                 "GaussianSmoothing", "DifferentiableBernoulliSampler", "random_eps", "norm", "SpectralNorm",
                 "CoordConv2d", "CSRN", "NOISE_DIM", "ENCODER_INPUT_FRAMES"):
        assert hasattr(m, name), name
    import coordconv, spatial_recurrent, spectral_normalization  # noqa: F401  (bare names, as main.py/models.py use)


def test_no_cpu_fallback(nets):
    """The product path must fail loudly instead of silently computing on the CPU."""
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    x = torch.zeros(1, 3, 3, 15, 19)
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        nets["encoder"](x)
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        nets["decoder"](torch.zeros(1, 16, 15, 19))


def test_interface_only_layers_match_golden():
    """CoordConv2d / CSRN (interface-only in the reference) against outputs recorded from the reference classes."""
    from scm_gan_b200.train_step import import_dropin_models
    import_dropin_models()
    import coordconv
    import spatial_recurrent
    g = torch.load(os.path.join(GOLDEN, "layers.pt"), weights_only=False)
    cc = coordconv.CoordConv2d(6 + 2, 16, 3, padding=1)
    cc.load_state_dict(g["coordconv"]["state"])
    torch.testing.assert_close(cc(g["coordconv"]["x"]), g["coordconv"]["y"], rtol=1e-5, atol=1e-5)
    cs = spatial_recurrent.CSRN(8)
    cs.load_state_dict(g["csrn"]["state"])
    with torch.no_grad():
        torch.testing.assert_close(cs(g["csrn"]["x"]), g["csrn"]["y"], rtol=1e-4, atol=1e-3)


# ---------------------------------------------------------------------------------------------------------------
# data-parallel gradient exchange on gloo, world_size = 2
# ---------------------------------------------------------------------------------------------------------------
class _ToyTrainer:
    NET_ORDER = ("reward_predictor", "encoder", "decoder", "transition")

    def __init__(self):
        torch.manual_seed(0)
        self.nets = {n: torch.nn.Linear(4, 3) for n in self.NET_ORDER}
        self.groups, self.net_of = [], []
        for ni, n in enumerate(self.NET_ORDER):
            for p in self.nets[n].parameters():
                self.groups.append((p, 0.1))
                self.net_of.append(ni)
                p.grad = torch.zeros_like(p)
        self.sync = None
        self.world_size = 1
        self.counts = {}
        self.counting = False
        for p, _ in self.groups:
            p.register_post_accumulate_grad_hook(self._on_grad)

    def _on_grad(self, p):
        if self.counting:
            self.counts[id(p)] = self.counts.get(id(p), 0) + 1
        elif self.sync is not None:
            self.sync.on_grad(p)

    def loss(self, x):
        # every net is used twice (like a 2-step rollout) except the encoder (once), transition unused if flag set
        h = self.nets["encoder"](x)
        out = 0
        for _ in range(2):
            out = out + self.nets["decoder"](x).sum() + self.nets["reward_predictor"](x).sum() * 0.5
            h = h + self.nets["transition"](x)
        return out + h.sum()


def _dp_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from scm_gan_b200.dp import BucketedGradSync
    t = _ToyTrainer()
    x = torch.arange(8, dtype=torch.float32).view(2, 4) * (rank + 1)
    # single-process reference gradients of this rank
    t.counting = True
    t.loss(x).backward()
    t.counting = False
    local = [p.grad.clone() for p, _ in t.groups]
    profile = dict(t.counts)
    sync = BucketedGradSync(t)
    assert t.world_size == world
    sync.zero()
    sync.arm(profile)
    t.loss(x).backward()
    sync.finish()
    got = [p.grad.clone() for p, _ in t.groups]
    # expected: sum over ranks of the local gradients
    exp = []
    for gl in local:
        e = gl.clone()
        dist.all_reduce(e)
        exp.append(e)
    ok = all(torch.allclose(a, b, rtol=1e-6, atol=1e-6) for a, b in zip(got, exp))
    views = all(p.grad.data_ptr() >= sync.buckets[sync._bucket_of[id(p)]]["flat"].data_ptr() for p, _ in t.groups)
    q.put((rank, ok and views))
    dist.destroy_process_group()


def test_bucketed_grad_sync_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_synthetic_env_contract():
    """MovingDots.make_env(): gym-style reset()/step() with the trajectory source's dynamics (for planner.play)."""
    import numpy as np
    from scm_gan_b200.synthetic import MovingDots
    src = MovingDots(3, 15, 19, 5, 2, seed=1)
    env = src.make_env()
    f = env.reset()
    assert f.shape == (3, 15, 19) and f.dtype == np.float32 and f[0].sum() == 1.0
    total, done, steps = 0.0, False, 0
    while not done:
        f, r, done, info = env.step(steps % 5)
        assert f.shape == (3, 15, 19) and r == sum(info.values())
        total += r
        steps += 1
    assert steps <= env.episode_length
    net_frame, rgb = src.convert_frame(f)
    assert net_frame.shape == (3, 15, 19) and rgb.shape == (15, 19, 3) and rgb.dtype == np.uint8


def test_planner_beam_table_matches_reference_enumeration():
    """planner.beam_actions builds the plan table of reference main.py:463-473 (every action pair + no-op tail)."""
    import numpy as np
    import torch
    from scm_gan_b200 import planner
    A, depth = 4, 12
    got = planner.beam_actions(A, lookahead=2, rollout_depth=depth, rollout_policy="noop")
    ref = []
    for i in range(A):
        for j in range(A):
            ref.append([i, j] + [0] * (depth - 2))
    assert got.dtype == torch.int64 and got.shape == (A * A, depth)
    assert np.array_equal(got.numpy(), np.asarray(ref))
    rnd = planner.beam_actions(A, rollout_depth=depth, rollout_policy="random", rng=np.random.RandomState(0))
    assert rnd.shape == (A * A, depth) and int(rnd.max()) < A and np.array_equal(rnd[:, :2].numpy(), np.asarray(ref)[:, :2])
    oh = planner.onehot(torch.tensor([0, 3]), A, "cpu")
    assert oh.shape == (2, A) and oh[1, 3] == 1 and oh.sum() == 2
    assert planner.onehot(2, A, "cpu").shape == (1, A)
