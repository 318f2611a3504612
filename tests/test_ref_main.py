"""The reference's own, unmodified main.py (LilJing/scm-gan) executed on the drop-in modules (VERDICT r1 item 9,
SURVEY.md section 8b/8c): `python main.py --env minipacman ...` through tests/ref_main_harness.py, which only supplies
the un-installable third-party packages (imutil, logutil, gym, gym_minipacman, matplotlib) and a synthetic MiniPacman
environment.  main.py is read from baseline/_ref (git-ignored verbatim copy of /root/reference, travels to the GPU box).

CPU: the harness itself is validated with the reference's own models.py (control arm).
GPU: training on the drop-ins makes the loss fall, every hot-path kernel launch goes through libscmgan.so, and the
     model-*.pth checkpoints main.py writes (main.py:133-141) load back through --load-from / --start-iter (79-90).
"""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS = os.path.join(ROOT, "tests", "ref_main_harness.py")
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_main_harness as H  # noqa: E402

pytestmark = pytest.mark.skipif(not H.prepare_ref_copy(), reason="no copy of the reference (baseline/_ref)")


def run(argv, cwd, env=None, timeout=900):
    e = dict(os.environ)
    e.pop("SCMGAN_HARNESS_MODELS", None)
    e.update(env or {})
    r = subprocess.run([sys.executable, HARNESS] + argv, cwd=cwd, env=e, capture_output=True, text=True,
                       timeout=timeout)
    assert r.returncode == 0, r.stdout[-3000:] + "\n" + r.stderr[-3000:]
    out = {}
    for line in r.stdout.splitlines():
        for tag in ("HARNESS_SERIES", "HARNESS_FULL", "HARNESS_MODELS", "HARNESS_LAUNCHES"):
            if line.startswith(tag + " "):
                out[tag] = line[len(tag) + 1:]
    assert "Finished execution, terminating" in r.stdout   # main.py:101
    return out, r.stdout


def test_harness_runs_reference_main_with_reference_models(tmp_path):
    """Control arm: main.py + the reference's own models.py on stock torch ops (CPU here)."""
    out, _ = run(["--env", "minipacman", "--train-iters", "4", "--batch-size", "2", "--horizon-max", "4"],
                 str(tmp_path), {"SCMGAN_HARNESS_MODELS": "reference", "CUDA_VISIBLE_DEVICES": ""})
    assert out["HARNESS_MODELS"].startswith(H.REF_COPY)
    series = json.loads(out["HARNESS_SERIES"])
    assert series["Reconstruction t=1"][2] == 4   # four iterations collected (main.py:196)


@pytest.mark.gpu
def test_reference_main_trains_on_dropin_modules(tmp_path):
    argv = ["--env", "minipacman", "--train-iters", "60", "--batch-size", "8", "--horizon-max", "6",
            "--enable-disentanglement-loss", "--enable-action-control-loss", "--counterfactual-horizon", "2"]
    out, stdout = run(argv, str(tmp_path), {"SCMGAN_HARNESS_SEED": "7"})
    assert out["HARNESS_MODELS"].startswith(H.DROPIN), out["HARNESS_MODELS"]
    assert int(out["HARNESS_LAUNCHES"].split()[0]) > 60 * 50   # the iterations ran on libscmgan.so kernels
    full = json.loads(out["HARNESS_FULL"])
    rec = full["Reconstruction t=1"]
    assert len(rec) == 60
    first, last = sum(rec[:5]) / 5, sum(rec[-5:]) / 5
    assert last < 0.8 * first, (first, last)
    assert all(v == v and abs(v) < 1e4 for s in full.values() for v in s)   # finite everywhere
    assert len(full["CF Disentanglement Loss"]) == 12 and len(full["CF Control Bias Loss"]) == 12   # every 5th (236)
    # horizon schedule of main.py:143-147: 3 + int(3*theta) -> t=2 terms appear from iteration 20, t=3 from 40
    assert len(full["Reconstruction t=2"]) == 41 and len(full["Reconstruction t=3"]) == 21


@pytest.mark.gpu
def test_reference_main_checkpoints_round_trip(tmp_path):
    a, b = tmp_path / "a", tmp_path / "b"
    a.mkdir(); b.mkdir()
    base = ["--env", "minipacman", "--batch-size", "4", "--horizon-max", "4"]
    # run 1: ITERS_PER_VIDEO lowered to 5 -> main.py:133-141 saves model-*.pth at iterations 5 and 10
    run(base + ["--train-iters", "10"], str(a), {"SCMGAN_HARNESS_ITERS_PER_VIDEO": "5", "SCMGAN_HARNESS_SEED": "3"})
    names = ["transition", "encoder", "decoder", "discriminator", "reward_predictor"]
    for n in names:
        assert (a / f"model-{n}.pth").is_file()
    # run 2: resume with --load-from / --start-iter; the harness saves what train() starts from
    dump = b / "entered.pt"
    out, _ = run(base + ["--train-iters", "14", "--start-iter", "11", "--load-from", str(a)], str(b),
                 {"SCMGAN_HARNESS_DUMP_STATE": str(dump), "SCMGAN_HARNESS_SEED": "4"})
    entered = torch.load(dump, map_location="cpu")
    for n in names:
        saved = torch.load(a / f"model-{n}.pth", map_location="cpu")
        assert set(saved) == set(entered[n])
        for k, v in saved.items():
            assert torch.equal(v, entered[n][k]), (n, k)
    series = json.loads(out["HARNESS_SERIES"])
    assert series["Reconstruction t=1"][2] == 4   # iterations 11..14
    # the key set is the reference's (weight_bar / weight_u / weight_v, unused bn_conv1.*), SURVEY.md section 5
    enc = torch.load(a / "model-encoder.pth", map_location="cpu")
    assert {"conv1.module.weight_bar", "conv1.module.weight_u", "conv1.module.weight_v", "bn_conv1.weight"} <= set(enc)
