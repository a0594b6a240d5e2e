#!/bin/bash
mkdir -p gpurun_out
python tools/prof_gemm.py > gpurun_out/plain_g.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'gemm_i8_tc' -s 12 -c 3 -o gpurun_out/prof_gemm -f python tools/prof_gemm.py > gpurun_out/ncu_g.log 2>&1
tail -3 gpurun_out/ncu_g.log
for st in 6 5 4 3; do QG_DBG_STAGES=$st python tools/gpu_perf.py --only stats_2sm_4096_f32 --out gpurun_out/perf_st$st.json 2>&1 | cut -c1-330; done
