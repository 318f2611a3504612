"""Per-kernel summary of an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...`).

  python profiles/summarize_launches.py gpurun_out/launches.csv [--out profiles/rNN_launches_summary.csv] [--grid]

Per-launch times under ncu are cold-cache and serialised (side-stream work counted in line): compare SHARES, not
absolutes.  --grid keeps launches of one kernel with different grid sizes apart (layer-level view).
"""
import argparse
import collections
import csv
import re


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--out", default=None)
    ap.add_argument("--grid", action="store_true")
    ap.add_argument("--title", default="")
    args = ap.parse_args()
    rows = []
    with open(args.csv, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"]
        name = re.sub(r"^void ", "", name)
        name = re.sub(r"\(.*\)$", "", name)
        if args.grid:
            name += " grid=" + r["Grid Size"].replace(" ", "")
        rows.append((name[:110], float(r["Metric Value"]) / 1e3))
    agg = collections.OrderedDict()
    for name, us in rows:
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
    total = sum(us for _, us in rows)
    out = [f"# {args.title}" if args.title else "# ncu launch list summary",
           f"# total {total / 1e3:.3f} ms over {len(rows)} launches; per-launch times are cold-cache and serialised under "
           f"ncu (side-stream work included): compare shares",
           "kernel,launches,total_ms,share_pct,avg_us"]
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f'"{name}",{n},{us / 1e3:.3f},{100 * us / total:.1f},{us / n:.1f}')
    text = "\n".join(out) + "\n"
    if args.out:
        with open(args.out, "w") as f:
            f.write(text)
    print(text)


if __name__ == "__main__":
    main()
