mkdir -p gpurun_out/r2c10
O=gpurun_out/r2c10
for lim in 0 74; do SRS_DIN_IMPL=rtp timeout -k 5 50 python profiles/trace_din_rt.py 4096 $lim > $O/trace_rtp_$lim.txt 2>&1; echo "trace $lim rc=$?"; done
tail -40 $O/trace_rtp_74.txt
