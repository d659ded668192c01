mkdir -p gpurun_out/r2c11
O=gpurun_out/r2c11
for v in default nowd nowd_lazy0 lazy1000 simple nowd_simple; do
  if [ $v = default ]; then unset SRS_CTR_LIB; else export SRS_CTR_LIB=$PWD/sparrowrecsys_b200/variants/libsrs_ctr_$v.so; fi
  SRS_DIN_IMPL=rtp timeout -k 5 50 python profiles/trace_din_rt.py 4096 74 > $O/trace_$v.txt 2>&1; echo "== $v trace rc=$?"
  grep -A12 "per-tile timeline" $O/trace_$v.txt | tail -6
  grep "^entry" $O/trace_$v.txt | tail -1 | cut -c1-900
  SRS_DIN_IMPL=rtp timeout -k 10 90 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --streams 2 > $O/bench_$v.json 2> $O/bench_$v.err; python -c "
import json; d=json.load(open('$O/bench_$v.json')); print('   bench S=2: %.1f M' % (d['value']/1e6))"
done
