#!/bin/bash
# First GPU calls of the next round, in order (each line is one `gpurun -- '<cmd>'`; wrap in the
# shown timeouts: the rth kernel has never run and a protocol error there is a hang).
#
# 1. the experimental half-SM DIN kernel against the oracle and against din_rt (14 cases)
#    SRS_TEST_RTH=1 timeout 150 python -m pytest tests/test_gpu_parity.py -k "rth and plain" -x -q   # plain mode first
#    SRS_TEST_RTH=1 timeout 150 python -m pytest tests/test_gpu_parity.py -k rth -x -q              # + SRS_DIN_RTH_BG=1
# 2. if green: its two co-residency modes against the current default, same box
#    for m in "" "SRS_DIN_IMPL=rth" "SRS_DIN_IMPL=rth SRS_DIN_RTH_CTAS=2" "SRS_DIN_IMPL=rth SRS_DIN_RTH_BG=1"; do
#      env $m timeout 60 python bench.py --steps 6000 --warmup 200 --no-cpu-baseline --streams 1
#      env $m timeout 60 python bench.py --steps 6000 --warmup 200 --no-cpu-baseline --streams 2 --sm-limit 0
#    done
#    (--streams 2 --sm-limit 0 with the default kernel is the control: full-width launches on two
#     streams cannot co-reside, they only hide the launch gap)
# 3. one `ncu --set full --clock-control none --import-source on -k regex:din_rth -c 1` capture of
#    the winner; summarise with profiles/summarize_ncu.py and profiles/hot_sass.py
# 4. the e2e leg: is it host-API bound now?  (184 M inf/s = 22 us per batch with 0.69 MB H2D =
#    16.8 us and 18.7 us of device time: count the API calls per batch in srs_predict_host_batches -
#    H2D, widen, forward, D2H, slot sync - and try 8 slots / one event per slot)
#    and SRS_ZERO_COPY_SCORES=1 (kernels write the scores into the caller's pinned buffer, no D2H call):
#      SRS_TEST_ZERO_COPY=1 python -m pytest tests/test_narrow_ids.py -k zero_copy -q
#      SRS_ZERO_COPY_SCORES=1 python bench.py --steps 6000 --warmup 200 --no-cpu-baseline
# 5. DIEN: dien_kernel runs one row per warp (70 us per 4096 rows at E = 10, T = 5); two or three
#    rows per warp (EP = 12 uses 12 of 32 lanes) is the obvious next step there.
echo "this file is a checklist, not a script to run as is"
