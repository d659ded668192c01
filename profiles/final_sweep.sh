#!/bin/bash
# End-of-round measurement pass for the headline workload (run on the GPU box through gpurun):
# full GPU test suite, default bench line, reference arm, phase trace, ncu launch list and one
# `ncu --set full` capture of the DIN kernel.  Outputs land in gpurun_out/.
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout -s KILL 300 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; cat gpurun_out/bench_final.json
timeout -s KILL 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>/dev/null; cat gpurun_out/bench_reference.json
timeout -s KILL 120 python profiles/trace_din_rt.py > gpurun_out/trace_final.txt 2>&1; tail -2 gpurun_out/trace_final.txt
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv \
    python bench.py --steps 50 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/b_launches.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:din_ -c 1 -s 3 -o gpurun_out/prof_din_final -f \
    python bench.py --steps 4 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/b_ncu_final.log 2>&1
ls -la gpurun_out/prof_din_final.ncu-rep gpurun_out/launches_final.csv
