N=${1:-2}
mkdir -p gpurun_out/r2c6
O=gpurun_out/r2c6
run() { timeout -k 10 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-e2e $2 > $O/bench_n${N}_$1.json 2> $O/bench_n${N}_$1.err; echo "bench $1 rc=$?"; }
run gather_fused "--gather fused"
run gather_fused_nograph "--gather fused --no-graph"
run gather_nccl "--gather nccl"
for f in $O/bench_n${N}_gather*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
  d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
  print(' value %.1f M ms/step %.3f  launch: %s' % (d['value']/1e6, d['ms_per_step'], d.get('detail',{}).get('launch','')[:80]))
except Exception as ex: print('ERR', ex)
PY
done
tail -n 4 $O/bench_n${N}_gather_fused.err | cut -c1-300
