"""Summarise ncu artefacts (run here, no GPU needed) into small tracked files under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches.csv           > profiles/x_launches.md
    python tools/ncu_summary.py report   gpurun_out/prof.ncu-rep [names..] > profiles/x_kernels.md
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warp_latency_per_inst_issued.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]


def launches(path):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[h]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = defaultdict(list)
    for r in rows[h + 1:]:
        if len(r) > vi:
            agg[r[ki]].append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    print("| kernel | launches | total ms | share | avg us |\n|---|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print("| `%s` | %d | %.3f | %.1f%% | %.1f |" % (k[:90], len(v), sum(v) / 1e6, 100 * sum(v) / tot,
                                                       sum(v) / len(v) / 1e3))


def report(path, names):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    data = rows[2:]
    names = names or ["launch %d" % i for i in range(len(data))]
    print("| metric | unit | " + " | ".join(names) + " |\n|---|---|" + "---|" * len(names))
    print("| kernel | | " + " | ".join("`%s`" % r[idx["Kernel Name"]][:48] for r in data) + " |")
    for k in KEYS:
        if k in idx:
            vals = []
            for r in data:
                try:
                    vals.append("%.4g" % float(r[idx[k]].replace(",", "")))
                except ValueError:
                    vals.append(r[idx[k]])
            print("| %s | %s | %s |" % (k, units[idx[k]], " | ".join(vals)))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        report(sys.argv[2], sys.argv[3:])
