mkdir -p gpurun_out/r2f1
O=gpurun_out/r2f1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt
timeout -k 10 900 python -m pytest tests -m gpu -q --timeout 300 > $O/gpu_suite.log 2>&1; echo "gpu suite rc=$?"; tail -6 $O/gpu_suite.log
timeout -k 10 200 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
timeout -k 10 300 python bench.py --steps 20 --warmup 5 > $O/bench_cfg3_din.json 2> $O/bench_cfg3_din.err; echo "bench default rc=$?"
timeout -k 10 300 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err; echo "reference rc=$?"
for w in cfg1_embeddingmlp cfg2_deepfm cfg2_deepfm_v2 cfg4_widendeep cfg4_neuralcf cfg4_twotowers ref_dien; do
  timeout -k 10 150 python bench.py --workload $w --steps 20 --warmup 5 --cpu-seconds 4 > $O/bench_$w.json 2> $O/bench_$w.err; echo "$w rc=$?"
done
timeout -k 10 300 python bench.py --workload cfg5_din --steps 20 --warmup 5 --cpu-seconds 4 > $O/bench_cfg5_din.json 2> $O/bench_cfg5_din.err; echo "cfg5 rc=$?"
timeout -k 10 300 python bench.py --workload cfg5_din --history zipf --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $O/bench_cfg5_din_zipf.json 2> $O/bench_cfg5_din_zipf.err; echo "cfg5 zipf rc=$?"
for b in 1024 65536; do
  timeout -k 10 300 python bench.py --workload cfg5_din --batch $b --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/bench_cfg5_din_b$b.json 2> $O/bench_cfg5_din_b$b.err; echo "cfg5 B=$b rc=$?"
done
SRS_DIN_IMPL=rtp timeout -k 10 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $O/bench_cfg3_din_rtp.json 2> $O/bench_cfg3_din_rtp.err; echo "rtp rc=$?"
# ncu: launch list of the default bench configuration (2 streams x 74 SMs, e2e legs with widen_u16), then one full capture per kernel
S="--dataset-batches 16 --host-batches 8 --steps 2 --warmup 1 --no-cpu-baseline"
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_cfg3_default.csv python bench.py $S > $O/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
cap() { # workload kernel-regex name
  timeout -k 10 200 ncu --set full --clock-control none --import-source on -k regex:$2 -s 24 -c 1 -o $O/ncu_$3 python bench.py --workload $1 $S --no-e2e > $O/ncu_$3.log 2>&1; echo "ncu $3 rc=$?"
}
cap cfg3_din din_rt_kernel cfg3_din_rt
cap cfg1_embeddingmlp embmlp_tc cfg1_embmlp_tc
cap cfg2_deepfm deepfm_tc cfg2_deepfm_tc
cap cfg2_deepfm_v2 deepfm2 cfg2_deepfm2
cap cfg4_widendeep embmlp_tc cfg4_widendeep
cap cfg4_neuralcf ncf_kernel cfg4_ncf
cap ref_dien dien_kernel ref_dien
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:din_rt64 -s 6 -c 1 -o $O/ncu_cfg5_din_rt64 python bench.py --workload cfg5_din --dataset-batches 4 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $O/ncu_cfg5.log 2>&1; echo "ncu cfg5 rc=$?"
ls -la $O/*.ncu-rep | awk '{print $5, $9}'
for f in $O/bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
  d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
  e=d.get('e2e',{}); e16=d.get('e2e_hist16',{})
  print(' value %.1f M  frac %s  e2e %.1f M  e2e16 %.1f M  kernel %s ms/step %.3f lat %s' % (d['value']/1e6, d.get('roofline',{}).get('frac'), e.get('value',0)/1e6, e16.get('value',0)/1e6, d.get('detail',{}).get('kernel'), d['ms_per_step'], (d.get('single_call_latency_us') or {}).get('median')))
  if 'cpu_baseline' in d: print('  cpu', round(d['cpu_baseline']['value']), d['cpu_baseline'].get('cores'))
except Exception as ex: print('ERR', ex)
PY
done
