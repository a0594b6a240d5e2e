#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 900 python tools/bench_configs.py > gpurun_out/configs.log 2>&1; echo "configs rc=$?"; tail -2 gpurun_out/configs.log | cut -c1-700
timeout 300 python tools/gpu_perf.py --only full_4096_pdl,stats_2sm_4096_f32 --out gpurun_out/perf_relu.json 2>&1 | grep -v "nosplit\|mnmajor" | cut -c1-250
