mkdir -p gpurun_out/r2c14
O=gpurun_out/r2c14
SRS_CTR_LIB=$PWD/sparrowrecsys_b200/variants/libsrs_ctr_tl.so SRS_DIN_IMPL=rtp timeout -k 5 50 python profiles/trace_din_rt.py 4096 74 > $O/trace_tl.txt 2>&1; echo "trace rc=$?"
grep -A14 "per-tile timeline" $O/trace_tl.txt | tail -9
