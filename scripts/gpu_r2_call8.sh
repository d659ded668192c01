mkdir -p gpurun_out/r2c9
O=gpurun_out/r2c9
timeout -k 5 200 python -m pytest tests/test_gpu_parity.py -k "rtp" -q -x --timeout 40 > $O/rtp_tests.log 2>&1; RTP=$?; echo "rtp tests rc=$RTP"; tail -8 $O/rtp_tests.log
if [ $RTP -eq 0 ]; then
  for lim in 0 74; do SRS_DIN_IMPL=rtp timeout -k 5 50 python profiles/trace_din_rt.py 4096 $lim > $O/trace_rtp_$lim.txt 2>&1; echo "trace $lim rc=$?"; tail -4 $O/trace_rtp_$lim.txt | head -1; done
  B="--steps 20 --warmup 5 --no-cpu-baseline --no-e2e"
  for S in 1 2 3 4 6; do
    SRS_DIN_IMPL=rtp timeout -k 10 90 python bench.py $B --streams $S > $O/bench_rtp_s$S.json 2> $O/bench_rtp_s$S.err; echo "rtp S=$S rc=$?"
  done
  SRS_DIN_IMPL=rtp timeout -k 10 240 ncu --set full --clock-control none --import-source on -k regex:din_rtp -s 6 -c 1 -o $O/ncu_rtp_lim74 python profiles/trace_din_rt.py 4096 74 > $O/ncu_rtp.log 2>&1; echo "ncu rtp rc=$?"
fi
for f in $O/bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
  d=json.load(open(sys.argv[1]))
  print(' value %.1f M  frac %s kernel %s ms/step %.3f' % (d['value']/1e6, d.get('roofline',{}).get('frac'), d.get('detail',{}).get('kernel'), d['ms_per_step']))
except Exception as ex: print('ERR', ex)
PY
done
