"""'Stock PyTorch on the same GPU' comparison: the oracle restatement of the reference step
(plain torch ops: cuDNN convolutions, eager launches) on the bench workload, fp32 with TF32 off and on, including
clip + torch.optim.Adam.  Reported in profiles/r01_notes.md; not part of bench.py.

    python profiles/stock_torch_gpu.py [iters]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from oracle import restated as R  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 5
dev = "cuda"
C, H, W, A, Rw, B, Hn = 3, 64, 64, 4, 1, 32, 10
T = Hn - 2
torch.backends.cudnn.benchmark = True
for tf32 in (False, True):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.allow_tf32 = tf32
    torch.manual_seed(0)
    nets = {"encoder": R.init_encoder(16, C), "decoder": R.init_decoder(16, C),
            "reward_predictor": R.init_reward_predictor(16, Rw), "transition": R.init_transition(16, A)}
    params = []
    for sd in nets.values():
        for k in list(sd):
            sd[k] = sd[k].to(dev)
            v = sd[k]
            if v.dtype.is_floating_point and not (k.endswith("_u") or k.endswith("_v") or "bn_conv1" in k):
                v.requires_grad_(True)
                params.append(v)
    opt = torch.optim.Adam(params, lr=1e-4)
    st, rw, dn, ac = R.synthetic_batch(B, Hn, C, H, W, A, Rw, seed=1234)
    st, rw, dn = st.to(dev), rw.to(dev), dn.to(dev)
    g = torch.Generator().manual_seed(1)
    times = []
    for it in range(iters + 2):
        cf_now = (it % 5 == 0)
        cf_idx, cf_perm = torch.randint(16, (B, 2), generator=g), torch.randperm(B, generator=g)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        opt.zero_grad()
        loss, _, _ = R.train_step_loss(nets, st, rw, dn, ac, num_actions=A, theta=1.0, enable_disentanglement=True,
                                       enable_action_control=True, cf_now=cf_now, counterfactual_horizon=3,
                                       cf_indices=cf_idx, cf_perm=cf_perm)
        loss.backward()
        torch.nn.utils.clip_grad_value_([p for p in params if p.grad is not None], 0.1)
        opt.step()
        torch.cuda.synchronize()
        if it >= 2:
            times.append(time.perf_counter() - t0)
    ms = 1e3 * sum(times) / len(times)
    print(f"stock torch (oracle restatement, eager, fp32, TF32 {'on' if tf32 else 'off'}): {ms:.1f} ms / iteration = "
          f"{B * T / ms * 1e3:.0f} frames/s over {len(times)} iterations", flush=True)
