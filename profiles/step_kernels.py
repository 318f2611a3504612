"""Kernel inventory of one training iteration (torch.profiler, eager launches): name, count, total device time.
    python profiles/step_kernels.py [--cf]      (a quick alternative to the ncu launch list)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
from scm_gan_b200 import synthetic as R  # synthetic batch generator
from scm_gan_b200.train_step import Trainer, build_nets

cf = "--cf" in sys.argv
dev = "cuda"
C, H, W, A, Rw, B, Hn = 3, 64, 64, 4, 1, 32, 10
nets = build_nets(C, A, Rw, seed=0)
for n in nets.values():
    n.train()
tr = Trainer(nets, loss_kwargs=dict(enable_disentanglement=True, enable_action_control=True, counterfactual_horizon=3))
st, rw, dn, ac = R.synthetic_batch(B, Hn, C, H, W, A, Rw, seed=1)
batch = {"states": st.to(dev), "rewards": rw.to(dev), "dones": dn.to(dev), "actions": torch.as_tensor(ac).to(dev),
         "cf_indices": torch.randint(16, (B, 2)).to(dev), "cf_perm": torch.randperm(B).to(dev)}
for _ in range(2):
    tr.step(batch, 1.0, cf_now=cf, use_graph=False)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    tr.step(batch, 1.0, cf_now=cf, use_graph=False)
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in rows)
n = sum(e.count for e in rows)
print(f"total device time {tot / 1e3:.3f} ms over {n} kernels/memops")
for e in rows[:40]:
    print(f"{e.device_time_total / 1e3:8.3f} ms {100 * e.device_time_total / tot:5.1f}% n={e.count:4d} avg={e.device_time_total / e.count:7.1f} us  {e.key[:100]}")

if "--seq" in sys.argv:
    # per-launch durations in launch order (our kernels only), to tell forward / dgrad / wgrad instances apart
    evs = [e for e in prof.events() if e.device_type.name == "CUDA" and "scm::" in e.name]
    evs.sort(key=lambda e: e.time_range.start)
    line = []
    for e in evs:
        short = e.name.split("scm::")[1].split("(")[0].replace("conv3x3_", "").replace("_kernel", "")
        line.append(f"{short}:{e.device_time:.0f}")
    print(" ".join(line))
