#!/bin/bash
# round 2, run 5 (1 GPU): L2 policy on the output stores; ncu --set full of the GEMM with fp32 / fp16 / int32 outputs and with the epilogue off
mkdir -p gpurun_out
timeout 600 python tools/gpu_perf.py --only r2_hint --out gpurun_out/r2_05_perf.json > gpurun_out/r2_05_perf.log 2>&1; echo "perf rc=$?"
python - <<'PY'
import json
p=json.load(open("gpurun_out/r2_05_perf.json"))
for k,v in p.items(): print(k, {a:(round(b,1) if isinstance(b,float) else b) for a,b in v.items() if a in ("us","us_with_stats","mma_total","mma_total_max","mma_wait_full","epi_total","error","first_start_to_last_end_us")})
PY
python tools/prof_gemm2.py 4096 > gpurun_out/r2_05_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_i8_tc -s 3 -c 3 -f -o gpurun_out/r2_gemm_full python tools/prof_gemm2.py 4096 > gpurun_out/r2_05_ncu.log 2>&1; echo "ncu rc=$?"
QG_DBG_NOEPI=1 python tools/prof_gemm2.py 4096 > gpurun_out/r2_05_plain2.log 2>&1 && \
QG_DBG_NOEPI=1 ncu --set full --clock-control none -k regex:gemm_i8_tc -s 3 -c 1 -f -o gpurun_out/r2_gemm_noepi python tools/prof_gemm2.py 4096 > gpurun_out/r2_05_ncu2.log 2>&1; echo "ncu2 rc=$?"
tail -3 gpurun_out/r2_05_ncu.log
