mkdir -p gpurun_out/r2c1
cd /root/repo
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2c1/smi.txt
SRS_TEST_RTH=1 timeout -k 10 400 python -m pytest tests/test_gpu_parity.py -k rth -q -x --timeout 60 > gpurun_out/r2c1/rth_tests.log 2>&1; echo "rth tests rc=$?"
tail -5 gpurun_out/r2c1/rth_tests.log
# continue past first failure to see which cases pass
SRS_TEST_RTH=1 timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -k rth -q --timeout 60 > gpurun_out/r2c1/rth_tests_all.log 2>&1; echo "rth tests(all) rc=$?"
tail -40 gpurun_out/r2c1/rth_tests_all.log
B="--steps 6000 --warmup 200 --no-cpu-baseline"
timeout -k 10 200 python bench.py $B > gpurun_out/r2c1/bench_rt_s2.json 2> gpurun_out/r2c1/bench_rt_s2.err; echo rc=$?
timeout -k 10 200 python bench.py $B --streams 1 > gpurun_out/r2c1/bench_rt_s1.json 2> gpurun_out/r2c1/bench_rt_s1.err; echo rc=$?
SRS_DIN_IMPL=rth SRS_DIN_RTH_CTAS=2 timeout -k 10 200 python bench.py $B --streams 1 > gpurun_out/r2c1/bench_rth_c2_s1.json 2> gpurun_out/r2c1/bench_rth_c2_s1.err; echo rc=$?
SRS_DIN_IMPL=rth SRS_DIN_RTH_CTAS=1 timeout -k 10 200 python bench.py $B --streams 2 --sm-limit 0 > gpurun_out/r2c1/bench_rth_c1_s2.json 2> gpurun_out/r2c1/bench_rth_c1_s2.err; echo rc=$?
SRS_DIN_IMPL=rth SRS_DIN_RTH_CTAS=2 timeout -k 10 200 python bench.py $B --streams 2 --sm-limit 0 > gpurun_out/r2c1/bench_rth_c2_s2.json 2> gpurun_out/r2c1/bench_rth_c2_s2.err; echo rc=$?
SRS_DIN_IMPL=rth SRS_DIN_RTH_CTAS=2 SRS_DIN_RTH_BG=1 timeout -k 10 200 python bench.py $B --streams 1 > gpurun_out/r2c1/bench_rth_c2_bg_s1.json 2> gpurun_out/r2c1/bench_rth_c2_bg_s1.err; echo rc=$?
for f in gpurun_out/r2c1/bench_*.json; do echo $f; python -c "
import json,sys
try:
  d=json.load(open('$f')); print(d['value']/1e6, d['e2e']['value']/1e6, d['roofline']['frac'], d['config']['kernel'])
except Exception as e: print('ERR',e)
"; done
tail -3 gpurun_out/r2c1/*.err
