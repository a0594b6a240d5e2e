#!/bin/bash
# round 2, run 2 (1 GPU): GEMM v2 (8 epilogue warps, per-warp scale slices, .read wait) -- parity subset, then the ablations again
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_outlier.py tests/test_gpu_reference.py -m gpu -x -q -p no:cacheprovider > gpurun_out/r2_02_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_02_pytest.log | cut -c1-300
timeout 900 python tools/gpu_perf.py --only r2_,lib_4096 --out gpurun_out/r2_02_perf.json > gpurun_out/r2_02_perf.log 2>&1; echo "perf rc=$?"
python - <<'PY'
import json
p=json.load(open("gpurun_out/r2_02_perf.json"))
for k,v in p.items(): print(k, {a:(round(b,1) if isinstance(b,float) else b) for a,b in v.items() if a in ("us","us_with_stats","mma_total","mma_total_max","mma_wait_full","epi_total","epi_wait_tfull","gemm_us","cols_us","rows_us","total_us","error","first_start_to_last_end_us","cublaslt_int8_us","cublas_fp16_us")})
PY
