N=${1:-2}
mkdir -p gpurun_out/r2c6
O=gpurun_out/r2c6
timeout -k 10 300 python -m pytest tests/test_multi_gpu.py -m gpu -q --timeout 300 > $O/mgpu_tests_${N}b.log 2>&1; echo "mgpu tests rc=$?"; tail -3 $O/mgpu_tests_${N}b.log
run() { timeout -k 10 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-e2e $2 > $O/bench_n${N}_$1.json 2> $O/bench_n${N}_$1.err; echo "bench $1 rc=$?"; }
run gather_fused "--gather fused"
for f in $O/bench_n${N}_gather_fused.json; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(' value %.1f M ms/step %.3f' % (d['value']/1e6, d['ms_per_step']))
PY
done
