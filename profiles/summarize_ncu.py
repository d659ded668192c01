#!/usr/bin/env python
"""Summarise ncu output brought back in gpurun_out/ into small tracked files.

  python profiles/summarize_ncu.py launches <launches.csv> <out.txt>
      per-kernel launch count / total / average device time from an
      `ncu --metrics gpu__time_duration.sum --csv` log
  python profiles/summarize_ncu.py full <prof.ncu-rep> <out.json> [kernel-regex]
      key metrics of the first matching launch of an `ncu --set full` capture
      (reads the report with `ncu -i ... --page raw --csv`)
"""
import collections
import csv
import json
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_tensor.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
    "smsp__warp_issue_stalled_wait_per_warp_active.pct",
    "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_membar_per_warp_active.pct",
    "smsp__warp_issue_stalled_sleeping_per_warp_active.pct",
]
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3,
        "s": 1e6}


def launches(path, out):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
    h = next(i for i, r in enumerate(rows) if r[0] == "ID")
    H = rows[h]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[h + 1:]:
        try:
            v = float(r[vi].replace(",", "")) * UNIT.get(r[ui], 1)
        except ValueError:
            continue
        agg.setdefault(r[ki], []).append(v)
    total = sum(sum(v) for v in agg.values())
    with open(out, "w") as f:
        f.write("# per-kernel device time from %s (cold-cache, serialised launches: compare shares)\n" % path)
        f.write("%-70s %6s %12s %10s %7s\n" % ("kernel", "n", "total_us", "avg_us", "share"))
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write("%-70s %6d %12.1f %10.2f %6.1f%%\n" % (k[:70], len(v), sum(v), sum(v) / len(v),
                                                           100 * sum(v) / total))
    print(open(out).read())


def full(rep, out, pattern=None):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    H, U = rows[0], rows[1]
    ki = H.index("Kernel Name")
    data = [r for r in rows[2:] if len(r) == len(H) and (not pattern or re.search(pattern, r[ki]))]
    r = data[0]
    res = {"report": rep, "kernel": r[ki], "launches_in_report": len(data)}
    for k in KEYS:
        if k in H:
            i = H.index(k)
            try:
                res[k] = float(r[i].replace(",", "")) * (UNIT.get(U[i], 1) if ("bytes" in k or "time" in k) else 1)
            except ValueError:
                res[k] = r[i]
    res["duration_us"] = res.get("gpu__time_duration.sum")
    res["dram_bytes_per_launch"] = res.get("dram__bytes_read.sum", 0) + res.get("dram__bytes_write.sum", 0)
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
