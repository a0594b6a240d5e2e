#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python tools/gpu_perf.py --only full_4096_pdl,full_8192,stats_2sm_4096_f32,stats_2sm_8192_f16 --out gpurun_out/perf_snake.json 2>&1 | grep -v "nosplit\|mnmajor" | cut -c1-250
timeout 900 python tools/bench_configs.py > gpurun_out/configs.log 2>&1; echo "configs rc=$?"; grep -E "opt66b|opt6.7b_fc2|transformer|mha" gpurun_out/configs.log | cut -c1-400
