# run with: gpurun --gpus 2 (or 4 / 8): multi-GPU tests and bench lines
N=${1:-2}
mkdir -p gpurun_out/r2c6
O=gpurun_out/r2c6
nvidia-smi topo -m > $O/topo_$N.txt 2>&1
timeout -k 10 600 python -m pytest tests/test_multi_gpu.py -m gpu -q --timeout 600 > $O/mgpu_tests_$N.log 2>&1; echo "mgpu tests rc=$?"; tail -15 $O/mgpu_tests_$N.log
run() { # name, extra args
  timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline $2 > $O/bench_n${N}_$1.json 2> $O/bench_n${N}_$1.err; echo "bench $1 rc=$?"
}
run plain "--no-e2e"
run gather_nccl "--gather nccl --no-e2e"
run gather_fused "--gather fused --no-e2e"
run plain_s1 "--streams 1 --no-e2e"
for f in $O/bench_n${N}_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
  d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
  e=d.get('e2e',{}); e16=d.get('e2e_hist16',{})
  print(' value %.1f M  e2e %.1f M  e2e16 %.1f M  kernel %s ms/step %.3f  %s' % (d['value']/1e6, e.get('value',0)/1e6, e16.get('value',0)/1e6, d.get('detail',{}).get('kernel'), d['ms_per_step'], d.get('detail',{}).get('parallelism','')[:120]))
except Exception as ex: print('ERR', ex)
PY
done
tail -n 5 $O/*.err $O/mgpu_tests_$N.log | cut -c1-300
