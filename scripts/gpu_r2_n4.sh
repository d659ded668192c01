# run with: gpurun --gpus 4 : multi-GPU test at world 4 and bench lines (priority order; tight timeouts)
N=${1:-4}
mkdir -p gpurun_out/r2n$N
O=gpurun_out/r2n$N
nvidia-smi topo -m > $O/topo_$N.txt 2>&1
timeout -k 10 150 python -m pytest tests/test_multi_gpu.py -m gpu -q --timeout 140 -k "$N" > $O/mgpu_tests_$N.log 2>&1; echo "mgpu tests rc=$?"; tail -3 $O/mgpu_tests_$N.log
run() { # name, extra args
  timeout -k 10 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline $2 > $O/bench_n${N}_$1.json 2> $O/bench_n${N}_$1.err; echo "bench $1 rc=$? $(grep -o '"value": [0-9.]*' $O/bench_n${N}_$1.json | head -3 | tr '\n' ' ')"
}
run gather_fused "--gather fused --no-e2e"
run plain ""
run gather_nccl "--gather nccl --no-e2e"
tail -n 3 $O/*.err | cut -c1-300
