#!/bin/bash
# round 2, run 4 (1 GPU): which part of the epilogue slows the main loop -- TMEM loads only / + staging / + stores / direct stores
mkdir -p gpurun_out
timeout 900 python tools/gpu_perf.py --only r2_epi,r2_stats_2sm_4096_f32_noepi --out gpurun_out/r2_04_perf.json > gpurun_out/r2_04_perf.log 2>&1; echo "perf rc=$?"
python - <<'PY'
import json
p=json.load(open("gpurun_out/r2_04_perf.json"))
for k,v in p.items(): print(k, {a:(round(b,1) if isinstance(b,float) else b) for a,b in v.items() if a in ("us","us_with_stats","mma_total","mma_total_max","mma_wait_full","epi_total","epi_wait_tfull","error","first_start_to_last_end_us")})
PY
