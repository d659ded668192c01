mkdir -p gpurun_out/r2f2
O=gpurun_out/r2f2
timeout -k 10 170 python bench.py --workload cfg5_din --steps 20 --warmup 5 --cpu-seconds 4 > $O/bench_cfg5_din.json 2> $O/bench_cfg5_din.err; rc=$?; echo "cfg5 rc=$rc"
tail -30 $O/bench_cfg5_din.err
if [ $rc -eq 0 ]; then
  for b in 1024 65536; do
    timeout -k 10 120 python bench.py --workload cfg5_din --batch $b --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/bench_cfg5_din_b$b.json 2> $O/bench_cfg5_din_b$b.err; echo "cfg5 B=$b rc=$?"
    tail -5 $O/bench_cfg5_din_b$b.err
  done
fi
cat $O/*.json | cut -c1-600
