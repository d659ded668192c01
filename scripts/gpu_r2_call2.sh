mkdir -p gpurun_out/r2c2
O=gpurun_out/r2c2
timeout -k 10 300 python -m pytest tests/test_gpu_parity.py -k "rtp" -x -q --timeout 60 > $O/rtp_tests.log 2>&1; echo "rtp tests rc=$?"
tail -15 $O/rtp_tests.log
for lim in 0 74 37; do SRS_DIN_IMPL=rtp timeout -k 5 60 python profiles/trace_din_rt.py 4096 $lim > $O/trace_rtp_$lim.txt 2>&1; echo "trace $lim rc=$?"; tail -4 $O/trace_rtp_$lim.txt; done
B="--steps 20 --warmup 5 --no-cpu-baseline --no-e2e"
for S in 1 2 4; do
  SRS_DIN_IMPL=rtp timeout -k 10 200 python bench.py $B --streams $S > $O/bench_rtp_s$S.json 2> $O/bench_rtp_s$S.err; echo "rtp S=$S rc=$?"
done
timeout -k 10 200 python bench.py $B --streams 2 > $O/bench_rt_s2.json 2> $O/bench_rt_s2.err; echo "rt S=2 rc=$?"
timeout -k 10 300 python bench.py --steps 20 --warmup 5 > $O/bench_default.json 2> $O/bench_default.err; echo "default rc=$?"
timeout -k 10 400 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err; echo "reference rc=$?"
for f in $O/bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
  d=json.load(open(sys.argv[1]))
  e=d.get('e2e',{}); e16=d.get('e2e_hist16',{})
  print(' value %.1f M  frac %s  e2e %.1f M  e2e16 %.1f M  kernel %s ms/step %.3f' % (d['value']/1e6, d.get('roofline',{}).get('frac'), e.get('value',0)/1e6, e16.get('value',0)/1e6, d.get('detail',{}).get('kernel'), d['ms_per_step']))
  if 'cpu_baseline' in d: print('  cpu', d['cpu_baseline']['value'], d['cpu_baseline'].get('cores'), d['cpu_baseline'].get('thread_sweep'))
  print('  lat', d.get('single_call_latency_us'), d.get('detail',{}).get('numa'))
except Exception as ex: print('ERR', ex)
PY
done
tail -n 3 $O/*.err
