#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 300 python tools/gpu_perf.py --only quant_4096,quant_8192,full_4096_pdl,full_8192,full_2048_pdl --out gpurun_out/perf_cols_rev.json 2>&1 | grep -v twopass | cut -c1-330
