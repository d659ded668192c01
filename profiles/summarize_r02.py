#!/usr/bin/env python
"""Round-2 evidence: turn the ncu reports / launch list brought back in gpurun_out/r2f1 into tracked files.

  profiles/r02/ncu_<name>.json        key metrics of one launch of each kernel (ncu --set full)
  profiles/ncu_bench_summary.json     {kernel name: {... dram_bytes_per_launch ...}}: bench.py's roofline.traffic
  profiles/r02/launches_cfg3_default.txt   per-kernel device-time shares of the default bench command
Run here (no GPU needed): python profiles/summarize_r02.py [gpurun_out/r2f1]"""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "profiles"))
import summarize_ncu as S   # noqa: E402

EXTRA = ["sm__icc_request_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
         "sm__inst_executed_pipe_tensor_op_hmma.sum", "smsp__pcsamp_warps_issue_stalled_long_scoreboard",
         "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_warps_issue_stalled_no_instructions",
         "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
         "smsp__pcsamp_warps_issue_stalled_lg_throttle", "smsp__pcsamp_warps_issue_stalled_mio_throttle",
         "smsp__pcsamp_warps_issue_stalled_selected"]
S.KEYS.extend(k for k in EXTRA if k not in S.KEYS)


# capture name -> bench workload (the captures ran `bench.py --workload W` at its default batch)
WORKLOAD_OF = {"ncu_cfg1_embmlp_tc": "cfg1_embeddingmlp", "ncu_cfg2_deepfm2": "cfg2_deepfm_v2",
               "ncu_cfg2_deepfm_tc": "cfg2_deepfm", "ncu_cfg3_din_rt": "cfg3_din", "ncu_cfg4_ncf": "cfg4_neuralcf",
               "ncu_cfg4_widendeep": "cfg4_widendeep", "ncu_cfg5_din_rt64": "cfg5_din", "ncu_ref_dien": "ref_dien"}
sys.path.insert(0, ROOT)
import bench   # noqa: E402
BENCH_BATCH = {w: v[0] for w, v in bench.WORKLOADS.items()}


def main():
    src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "r2f1")
    out_dir = os.path.join(ROOT, "profiles", "r02")
    os.makedirs(out_dir, exist_ok=True)
    summary = {}
    for rep in sorted(glob.glob(os.path.join(src, "ncu_*.ncu-rep"))):
        name = os.path.basename(rep)[:-8]
        out = os.path.join(out_dir, name + ".json")
        try:
            S.full(rep, out)
        except Exception as e:           # a capture that matched no launch
            print("skip", rep, e)
            continue
        d = json.load(open(out))
        d["report"] = os.path.relpath(rep, ROOT)
        json.dump(d, open(out, "w"), indent=1)
        kernel = d["kernel"].split("(")[0].split("<")[0].split("::")[-1].replace("void ", "").strip()
        workload = WORKLOAD_OF[name]
        summary[workload] = dict(kernel=kernel, batch=BENCH_BATCH[workload], **{k: d.get(k) for k in (
            "duration_us", "dram_bytes_per_launch", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "launch__grid_size", "launch__block_size", "report")})
    json.dump(summary, open(os.path.join(ROOT, "profiles", "ncu_bench_summary.json"), "w"), indent=1)
    lc = os.path.join(src, "launches_cfg3_default.csv")
    if os.path.exists(lc):
        S.launches(lc, os.path.join(out_dir, "launches_cfg3_default.txt"))
    for f in glob.glob(os.path.join(src, "bench_*.json")) + glob.glob(os.path.join(src, "*.log")) + \
            glob.glob(os.path.join(src, "smi.txt")):
        if 0 < os.path.getsize(f) < 300000:
            subprocess.call(["cp", f, os.path.join(out_dir, os.path.basename(f))])


if __name__ == "__main__":
    main()
