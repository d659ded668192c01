mkdir -p gpurun_out/r2c12
O=gpurun_out/r2c12
for v in nowd_simple; do
  export SRS_CTR_LIB=$PWD/sparrowrecsys_b200/variants/libsrs_ctr_$v.so
  SRS_DIN_IMPL=rtp timeout -k 5 50 python profiles/trace_din_rt.py 4096 74 > $O/trace_$v.txt 2>&1; echo "== $v trace rc=$?"
  grep -A14 "per-tile timeline" $O/trace_$v.txt | tail -9
done
