#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
for c in 8 14 4 16 6; do echo "== QG_COLS_CTAS_PER_SM=$c"; QG_COLS_CTAS_PER_SM=$c timeout 300 python tools/gpu_perf.py --only quant_4096,quant_8192,full_4096_pdl --out gpurun_out/perf_cols$c.json 2>&1 | grep -v twopass | cut -c1-330; done
