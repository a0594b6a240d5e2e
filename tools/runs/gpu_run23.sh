#!/bin/bash
# launch list of the bench command with the caches left as the pipeline leaves them (single-pass metrics, no replay):
# does the column quantizer's second pass find W in L2?
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_launches.log | cut -c1-300
timeout 300 python tools/gpu_perf.py --only quant_4096,full_4096_pdl,full_8192 --out gpurun_out/perf_cols_final.json 2>&1 | grep -v twopass | cut -c1-330
