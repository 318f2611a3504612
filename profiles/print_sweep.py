"""Table view of profiles/sweep.py output (one JSON line per configuration)."""
import json
import sys

for path in sys.argv[1:]:
    for line in open(path):
        if not line.startswith("{"):
            continue
        d = json.loads(line)
        c = d["config"]
        print(f"N={d['n_gpus']} {c['workload']:10s} B/GPU={c['batch_per_gpu']:5d} T={c['T']:2d} cf={int(c['cf'])} "
              f"{d['ms_per_step']:9.3f} ms {d['value']:11.0f} frames/s {100 * d['frac_of_sustained_bf16_peak']:5.1f} % of "
              f"sustained peak  mem {d['peak_mem_gib']:6.1f} GiB  {d['scaling']:6s} {c['note']}")
