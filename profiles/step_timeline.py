"""Timeline of ONE CUDA-graph replay of the training iteration (CUPTI kernel records through torch.profiler): where the
9 ms go that are not the sum of the kernels' stand-alone times.  Prints, for the compute stream, busy time, idle gaps and
which kernel pairs the gaps sit between; for the side stream(s) the same busy time.

    python profiles/step_timeline.py [--cf] [--json gpurun_out/r02_timeline.json]
"""
import argparse
import collections
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from scm_gan_b200 import synthetic as R  # noqa: E402
from scm_gan_b200.train_step import Trainer, build_nets  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cf", action="store_true")
ap.add_argument("--json", default=None)
args = ap.parse_args()
dev = "cuda"
C, H, W, A, Rw, B, Hn = 3, 64, 64, 4, 1, 32, 10
nets = build_nets(C, A, Rw, seed=0)
for n in nets.values():
    n.train()
tr = Trainer(nets, loss_kwargs=dict(enable_disentanglement=True, enable_action_control=True, counterfactual_horizon=3))
st, rw, dn, ac = R.synthetic_batch(B, Hn, C, H, W, A, Rw, seed=1)
batch = {"states": st.to(dev), "rewards": rw.to(dev), "dones": dn.to(dev), "actions": torch.as_tensor(ac).to(dev),
         "cf_indices": torch.randint(16, (B, 2)).to(dev), "cf_perm": torch.randperm(B).to(dev)}
for _ in range(4):
    tr.step(batch, 1.0, cf_now=args.cf, use_graph=True)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        tr.step(batch, 1.0, cf_now=args.cf, use_graph=True)
    torch.cuda.synchronize()
ev = []
for e in prof.events():
    if e.device_type is not None and "cuda" in str(e.device_type).lower() and e.time_range is not None:
        name = e.name
        if name.startswith("Memcpy") or name.startswith("Memset") or "(" in name or "kernel" in name.lower() or "nccl" in name.lower():
            ev.append((e.time_range.start, e.time_range.end, name, getattr(e, "device_resource_id", None) or 0))
ev.sort()
if not ev:
    sys.exit("no kernel records (CUPTI does not trace graph nodes here)")
# one replay = a third of the records; take the middle one
n = len(ev) // 3
one = ev[n:2 * n]
t0, t1 = one[0][0], max(e[1] for e in one)
streams = collections.defaultdict(list)
for s, e, name, sid in one:
    streams[sid].append((s, e, name))
main_sid = max(streams, key=lambda k: sum(e - s for s, e, _ in streams[k]))
print(f"replay span {(t1 - t0) / 1e3:.3f} ms, {len(one)} records, streams: "
      + ", ".join(f"{sid}: {len(v)} kernels, busy {sum(e - s for s, e, _ in v) / 1e3:.3f} ms" for sid, v in streams.items()))


def short(nm):
    nm = nm.replace("scm::", "").replace("void ", "")
    return nm.split("(")[0][:48]


m = sorted(streams[main_sid])
gaps = collections.Counter()
gapn = collections.Counter()
idle = 0.0
for (s0, e0, n0), (s1, e1, n1) in zip(m, m[1:]):
    g = s1 - e0
    if g > 0:
        idle += g
        gaps[(short(n0), short(n1))] += g
        gapn[(short(n0), short(n1))] += 1
print(f"compute stream: busy {sum(e - s for s, e, _ in m) / 1e3:.3f} ms, idle between kernels {idle / 1e3:.3f} ms "
      f"over {len(m) - 1} boundaries ({idle / max(1, len(m) - 1):.2f} us each on average)")
print("largest idle contributors (after kernel -> before kernel: total us, count):")
for k, v in gaps.most_common(14):
    print(f"  {k[0]:>48s} -> {k[1]:<48s} {v:8.1f} us  x{gapn[k]}")
dur = collections.Counter()
cnt = collections.Counter()
for s, e, nm in m:
    dur[short(nm)] += e - s
    cnt[short(nm)] += 1
print("compute-stream kernel time in the replay (us total, count, us each):")
for k, v in dur.most_common(16):
    print(f"  {k:<50s} {v:9.1f}  x{cnt[k]:<4d} {v / cnt[k]:7.1f}")
if args.json:
    json.dump({"span_ms": (t1 - t0) / 1e3, "streams": {str(k): [(s - t0, e - t0, short(nm)) for s, e, nm in v]
                                                         for k, v in streams.items()}}, open(args.json, "w"))
