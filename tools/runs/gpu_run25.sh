#!/bin/bash
mkdir -p gpurun_out
for lib in libqgemm.so libqgemm_l2hint.so libqgemm.so libqgemm_l2hint.so; do echo "== $lib"; QG_LIB=$lib timeout 300 python tools/gpu_perf.py --only quant_4096,full_4096_pdl --out gpurun_out/perf_$lib.json 2>&1 | grep -v twopass | cut -c1-330; done
QG_LIB=libqgemm_l2hint.so python tools/prof_quant.py > gpurun_out/plain_q.log 2>&1 &&
QG_LIB=libqgemm_l2hint.so ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct -k regex:'absmax_cols_partial|quant_cols_kernel' -c 12 --csv --log-file gpurun_out/launches_l2hint.csv python tools/prof_quant.py > gpurun_out/ncu_q.log 2>&1
grep -E "absmax_cols|quant_cols" gpurun_out/launches_l2hint.csv | awk -F'","' '{print $5, $(NF-2), $(NF)}' | cut -c1-160 | tail -16
