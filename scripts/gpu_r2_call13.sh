mkdir -p gpurun_out/r2c16
O=gpurun_out/r2c16
timeout -k 5 200 python -m pytest tests/test_gpu_parity.py -k "rtp" -q -x --timeout 40 > $O/rtp_tests.log 2>&1; RTP=$?; echo "rtp tests rc=$RTP"; tail -8 $O/rtp_tests.log
if [ $RTP -eq 0 ]; then
  SRS_CTR_LIB=$PWD/sparrowrecsys_b200/variants/libsrs_ctr_tl.so SRS_DIN_IMPL=rtp timeout -k 5 50 python profiles/trace_din_rt.py 4096 74 > $O/trace_tl.txt 2>&1; echo "trace rc=$?"
  grep "^entry" $O/trace_tl.txt | tail -1 | cut -c1-1000
  grep -A14 "per-tile timeline" $O/trace_tl.txt | tail -9
  for S in 1 2 4; do
    SRS_DIN_IMPL=rtp timeout -k 10 90 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --streams $S > $O/bench_rtp_s$S.json 2> $O/bench_rtp_s$S.err; python -c "
import json; d=json.load(open('$O/bench_rtp_s$S.json')); print('   bench S=$S: %.1f M' % (d['value']/1e6))"
  done
fi
