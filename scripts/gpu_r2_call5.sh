mkdir -p gpurun_out/r2c5
O=gpurun_out/r2c5
timeout -k 5 60 python -m pytest tests/test_gpu_parity.py -k "test_din_rtp_kernel and 32-50-28-0" -x -q --timeout 30 > $O/rtp_first.log 2>&1; RTP=$?
echo "rtp first case rc=$RTP"; tail -15 $O/rtp_first.log
if [ $RTP -eq 0 ]; then
  timeout -k 5 300 python -m pytest tests/test_gpu_parity.py -k "rtp" -q --timeout 40 > $O/rtp_tests.log 2>&1; echo "rtp tests rc=$?"; tail -30 $O/rtp_tests.log
  for lim in 0 74 37; do SRS_DIN_IMPL=rtp timeout -k 5 50 python profiles/trace_din_rt.py 4096 $lim > $O/trace_rtp_$lim.txt 2>&1; echo "trace $lim rc=$?"; tail -4 $O/trace_rtp_$lim.txt; done
  B="--steps 20 --warmup 5 --no-cpu-baseline --no-e2e"
  for S in 1 2 3 4; do
    SRS_DIN_IMPL=rtp timeout -k 10 90 python bench.py $B --streams $S > $O/bench_rtp_s$S.json 2> $O/bench_rtp_s$S.err; echo "rtp S=$S rc=$?"
  done
fi
timeout -k 10 600 python -m pytest tests/test_featurestore.py tests/test_ranking.py tests/test_narrow_ids.py -m gpu -q --timeout 120 > $O/gpu_some.log 2>&1; echo "gpu subset rc=$?"; tail -30 $O/gpu_some.log
timeout -k 10 200 python profiles/rank_latency_r2.py > $O/rank_latency.log 2>&1; echo "rank latency rc=$?"; tail -3 $O/rank_latency.log | cut -c1-1800; cp gpurun_out/rank_latency_r02.json $O/ 2>/dev/null
for f in $O/bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
  d=json.load(open(sys.argv[1]))
  print(' value %.1f M  frac %s kernel %s ms/step %.3f' % (d['value']/1e6, d.get('roofline',{}).get('frac'), d.get('detail',{}).get('kernel'), d['ms_per_step']))
except Exception as ex: print('ERR', ex)
PY
done
tail -n 3 $O/*.err
