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

# Later in the round (second session), one gpurun call each:
#   python -m pytest tests/test_ranking.py -m gpu -x -q; python profiles/rank_latency.py
#   python -m pytest tests/test_seq_dien.py -m gpu -x -q; python bench.py --workload ref_dien --steps 2000 --warmup 20 --cpu-seconds 3;
#       python -m pytest tests -m gpu -x -q --deselect tests/test_seq_dien.py
#   python -m pytest tests/test_narrow_ids.py tests/test_ranking.py -m gpu -q;
#       for S in 2 4; do python bench.py --steps 6000 --warmup 200 --no-cpu-baseline --narrow-ids auto --streams $S; done;
#       python bench.py --steps 6000 --warmup 200 --no-cpu-baseline --streams 1 --narrow-ids off
# (outputs: rank_*_r01.*, dien_tests_r01.log, gpu_suite_r01_session2.log, bench_r01/bench_ref_dien.json,
#  narrow_ids_sm_limit_tests_r01.log, bench_r01_streams/)
