#!/bin/bash
mkdir -p gpurun_out
python tools/prof_gemm.py > gpurun_out/plain_g.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'gemm_i8_tc' -s 4 -c 4 -o gpurun_out/prof_gemm -f python tools/prof_gemm.py > gpurun_out/ncu_g.log 2>&1
tail -2 gpurun_out/ncu_g.log
python tools/gpu_perf.py --only stats_2sm_4096_f32,stats_1sm_4096_f32,stats_2sm_8192_f16 --out gpurun_out/perf_base.json 2>&1 | cut -c1-420
QG_DBG_NOLOAD=1 python tools/gpu_perf.py --only stats_2sm_4096_f32,stats_1sm_4096_f32,stats_2sm_8192_f16 --out gpurun_out/perf_noload.json 2>&1 | cut -c1-420
QG_DBG_ALL_HALF=1 python tools/gpu_perf.py --only stats_2sm_4096_f32,stats_1sm_4096_f32 --out gpurun_out/perf_allhalf.json 2>&1 | cut -c1-420
QG_DBG_ALL_HALF=1 QG_DBG_NOLOAD=1 python tools/gpu_perf.py --only stats_2sm_4096_f32,stats_1sm_4096_f32 --out gpurun_out/perf_allhalf_noload.json 2>&1 | cut -c1-420
