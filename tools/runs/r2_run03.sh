#!/bin/bash
# round 2, run 3 (1 GPU): producer / MMA warps moved to the highest warp ids -- GEMM parity subset + ablations
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider -k "gemm or dequant or quantized_mm or linear" > gpurun_out/r2_03_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_03_pytest.log | cut -c1-300
timeout 900 python tools/gpu_perf.py --only r2_stats_2sm_4096,r2_stats_2sm_8192,r2_inop,r2_clusters8_full,full_4096_pdl,full_8192 --out gpurun_out/r2_03_perf.json > gpurun_out/r2_03_perf.log 2>&1; echo "perf rc=$?"
python - <<'PY'
import json
p=json.load(open("gpurun_out/r2_03_perf.json"))
for k,v in p.items(): print(k, {a:(round(b,1) if isinstance(b,float) else b) for a,b in v.items() if a in ("us","us_with_stats","mma_total","mma_total_max","mma_wait_full","epi_total","epi_wait_tfull","gemm_us","cols_us","rows_us","total_us","error","first_start_to_last_end_us","full_us","full_graph_us","linear_cached_w_us")})
PY
