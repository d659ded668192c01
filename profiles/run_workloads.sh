#!/bin/bash
# Run bench.py on every BASELINE workload (1 GPU) and keep the JSON lines: run on the GPU box.
mkdir -p gpurun_out
for w in cfg1_embeddingmlp cfg2_deepfm cfg2_deepfm_v2 cfg3_din cfg4_widendeep cfg4_neuralcf cfg4_twotowers cfg5_din; do
  steps=3000; [ "$w" = "cfg5_din" ] && steps=300
  timeout -k 10 400 python bench.py --workload $w --steps $steps --warmup 30 --cpu-seconds 4 \
      > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err || echo "FAILED $w"
  tail -c 600 gpurun_out/bench_$w.err
  python - "$w" <<'PY'
import json, sys
w = sys.argv[1]
try:
    d = json.loads(open("gpurun_out/bench_%s.json" % w).read().strip().splitlines()[-1])
    print("%-20s value %.4g inf/s  %.2f us/step  e2e %.4g  frac %.3f  cpu %.4g  kernel %s" % (
        w, d["value"], 1e3 * d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"],
        d.get("cpu_baseline", {}).get("value", float("nan")), d["config"]["kernel"]))
except Exception as e:
    print(w, "no result", e)
PY
done
