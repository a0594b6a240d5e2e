#!/bin/bash
mkdir -p gpurun_out
for st in 6 4 3; do QG_DBG_STAGES=$st python tools/gpu_perf.py --only stats_2sm_4096_f32,stats_2sm_8192_f16 --out gpurun_out/perf_st$st.json 2>&1 | cut -c1-300; done
QG_DBG_NOLOAD=1 python tools/gpu_perf.py --only stats_2sm_4096_f32,stats_2sm_8192_f16 --out gpurun_out/perf_noload.json 2>&1 | cut -c1-300
QG_PERCALL_KMAJOR=0 python tools/gpu_perf.py --only full_4096_pdl,full_8192,full_2048_pdl --out gpurun_out/perf_mn.json 2>&1 | cut -c1-300
QG_PERCALL_KMAJOR=1 python tools/gpu_perf.py --only full_4096_pdl,full_8192,full_2048_pdl --out gpurun_out/perf_k.json 2>&1 | cut -c1-300
