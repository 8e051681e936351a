#!/usr/bin/env python
"""Turn the ncu outputs brought back in gpurun_out/ into the small committed
summaries under profiles/ (launch shares of the bench command, key metrics and
top stall reasons of the --set full captures).  Run here (no GPU needed)."""
import collections
import csv
import json
import subprocess
import sys

TAG = sys.argv[1] if len(sys.argv) > 1 else "r1"


def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H, data = rows[hdr], rows[hdr + 1:]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        name = r[ki].split("(")[0][:100]
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)  # -> us
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    lines = ["# ncu launch list of: python bench.py --steps 2 --warmup 3 --no-extra  (gpu__time_duration.sum, --clock-control none)",
             "# per-launch times are cold-cache and serialised: compare SHARES, not absolutes.  The at::* kernels are",
             "# torch's synthetic-data generation and result comparison OUTSIDE the timed region; inside it only",
             "# k_fused_reduce (+ one memset per launch) runs.",
             f"# {len(data)} launches, total {tot / 1e3:.2f} ms", "kernel,launches,total_us,share,avg_us"]
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"\"{n}\",{c},{t:.1f},{t / tot:.4f},{t / c:.1f}")
    open(out, "w").write("\n".join(lines) + "\n")
    return lines


def full(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    d, units = dict(zip(rows[0], rows[-1])), dict(zip(rows[0], rows[1]))
    keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
            "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "smsp__warps_eligible.avg.per_cycle_active", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size"]
    s = {k: [d.get(k), units.get(k)] for k in keys}
    stalls = sorted(((float(v), k) for k, v in d.items()
                     if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and "not_issued" not in k
                     and v not in (None, "")), reverse=True)
    s["top_stalls_warps_per_issue"] = {
        k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): round(v, 3)
        for v, k in stalls[:7]}
    for k in ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
              "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"):
        s[k] = [d.get(k), units.get(k)]
    return s


if __name__ == "__main__":
    import glob
    import os

    SUF = sys.argv[2] if len(sys.argv) > 2 else TAG
    src = f"gpurun_out/launches_bench_{SUF}.csv"
    if os.path.exists(src):
        for ln in launches(src, f"profiles/{TAG}_launches_bench.csv")[:10]:
            print(ln)
    summ = {}
    names = {"small": "200x200x8760_100shapes", "big": "1440x720x432_3000shapes"}
    for rep in sorted(glob.glob(f"gpurun_out/prof_*_{SUF}.ncu-rep")):
        kind, size = os.path.basename(rep).split("_")[1:3]
        try:
            summ[f"{kind}_{names.get(size, size)}"] = full(rep)
        except Exception as e:  # noqa: BLE001
            print("skip", rep, e)
    json.dump(summ, open(f"profiles/{TAG}_ncu_full_summary.json", "w"), indent=1)
    for tag, s in summ.items():
        print("==", tag, s["Kernel Name"][0][:70])
        for k in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
                  "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
                  "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
                  "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"):
            print("  ", k, s.get(k))
        print("   stalls", s["top_stalls_warps_per_issue"])
