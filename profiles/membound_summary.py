"""HBM evidence for the memory- / latency-bound kernels of the step.

Input: the ncu pass  `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum
--clock-control none -k regex:... python bench.py --steps 1 --warmup 3 --no-cpu-baseline --cf-phase 1`  (pong64 bench
workload, B = 32, 64x64, T = 8).  Output: one row per (kernel, grid): launches, mean duration, measured DRAM bytes, L2
bytes, ALGORITHMIC bytes (unique tensor bytes in + out) and the fractions of the measured HBM peak
(MEASURED_PEAKS.json copy bandwidth, 6546.6 GB/s) they amount to.  ncu serialises launches and runs them cold, so the
durations are upper bounds for the step; DRAM writes are mostly absent because outputs of this size stay in the 126 MB
L2 until evicted.

    python profiles/membound_summary.py gpurun_out/r02_membound.csv > profiles/r02_membound_kernels.csv
"""
import collections
import csv
import sys

HBM = 6546.6e9
B, H, W, T, L, C = 32, 64, 64, 8, 16, 3
HW, PL = H * W, (H + 2) * (W + 2)
P_TOTAL = 1_166_710  # trainable floats

# (kernel, grid) -> (algorithmic bytes, what) at the bench shapes; None = weight-sized, L2-resident working set
ALG = {
    ("pack_nchw_to_plane_kernel", "(1089, 1, 1)"): (B * L * HW * 4 + B * PL * 16 * 2, "z fp32 NCHW -> 16-ch plane (B=32)"),
    ("pack_nchw_to_plane_kernel", "(8712, 1, 1)"): (B * T * L * HW * 4 + B * T * PL * 16 * 2, "latents of all T steps -> plane (B*T=256)"),
    ("bce_logits_kernel", "(4, 256, 1)"): (B * T * C * HW * (4 + 4 + 4), "logits + target read, dlogits written (B*T=256)"),
    ("clip_adam_kernel", "(64, 28, 1)"): (P_TOTAL * 28, "p, g, m, v read; p, m, v written"),
    ("plane_colsum_kernel", "(16, 32, 1)"): (B * PL * 16 * 2, "16-ch gradient plane read (bias gradient)"),
    ("plane_colsum_kernel", "(2, 256, 1)"): (B * T * PL * 16 * 2, "16-ch gradient plane read (B*T=256)"),
    ("philox_fill_kernel", "(2048, 1, 1)"): (B * L * HW * 4, "uniforms written"),
    ("reward_head_fwd_kernel", "(256, 1, 1)"): (B * T * 6 * 30 * 30 * 4, "conv2 lattice read"),
    ("reward_head_bwd_kernel", "(8712, 1, 1)"): (B * T * PL * 16 * 2, "gradient plane written"),
    ("wgrad_reduce_kernel<8>", "(1153, 1, 1)"): (None, "split-K partials of a 128x128 layer (28 MB) -> 0.6 MB gradient"),
    ("wgrad_reduce_kernel<8>", "(1152, 1, 1)"): (None, "same, without bias"),
}


def main():
    path = sys.argv[1]
    rows = list(csv.DictReader([ln for ln in open(path) if ln.startswith('"')]))
    agg = collections.OrderedDict()
    for r in rows:
        name = r["Kernel Name"].split("(")[0].replace("void ", "").replace("scm::", "")
        key = (name, r["Grid Size"])
        agg.setdefault(key, collections.defaultdict(list))[r["Metric Name"]].append(float(r["Metric Value"].replace(",", "")))
    print("# memory- / latency-bound kernels of one pong64 training iteration (ncu, cold, serialised); HBM peak 6546.6 GB/s")
    print("kernel,grid,launches,mean_us,dram_read_MB,dram_write_MB,l2_MB,dram_GBps,dram_frac_of_peak,alg_MB,alg_GBps,"
          "alg_frac_of_peak,bound,what")
    for (name, grid), v in sorted(agg.items(), key=lambda kv: -sum(kv[1]["gpu__time_duration.sum"])):
        n = len(v["gpu__time_duration.sum"])
        t = sum(v["gpu__time_duration.sum"]) / n * 1e-9
        rd, wr = sum(v["dram__bytes_read.sum"]) / n, sum(v["dram__bytes_write.sum"]) / n
        l2 = sum(v["lts__t_bytes.sum"]) / n
        alg, what = ALG.get((name, grid), (None, ""))
        dram_bw = (rd + wr) / t
        bound = "hbm" if dram_bw / HBM > 0.5 else ("latency" if t < 30e-6 else "l2/latency")
        print(f"\"{name}\",\"{grid}\",{n},{t * 1e6:.1f},{rd / 1e6:.2f},{wr / 1e6:.2f},{l2 / 1e6:.1f},{dram_bw / 1e9:.0f},"
              f"{dram_bw / HBM:.3f},{'' if alg is None else f'{alg / 1e6:.2f}'},"
              f"{'' if alg is None else f'{alg / t / 1e9:.0f}'},{'' if alg is None else f'{alg / t / HBM:.3f}'},{bound},"
              f"\"{what}\"")


if __name__ == "__main__":
    main()
