mkdir -p gpurun_out/r2c17
O=gpurun_out/r2c17
for v in ahead1 ahead0; do
  export SRS_CTR_LIB=$PWD/sparrowrecsys_b200/variants/libsrs_ctr_$v.so
  timeout -k 5 100 python -m pytest tests/test_gpu_parity.py -k "rtp and (32-50-4096-74 or 32-50-1500-3 or 32-9-100-1)" -q -x --timeout 40 > $O/tests_$v.log 2>&1; echo "$v tests rc=$?"
  for S in 1 2 4; do
    SRS_DIN_IMPL=rtp timeout -k 10 90 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --streams $S > $O/bench_${v}_s$S.json 2> $O/bench_${v}_s$S.err; python -c "
import json; d=json.load(open('$O/bench_${v}_s$S.json')); print('   $v bench S=$S: %.1f M' % (d['value']/1e6))"
  done
done
