#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_attention.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python tools/bench_configs.py > gpurun_out/configs.log 2>&1; echo "configs rc=$?"; tail -16 gpurun_out/configs.log | cut -c1-700
