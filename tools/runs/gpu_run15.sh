#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log | cut -c1-300
python tools/gpu_perf.py --only lib_4096,lib_8192,stats_,full_4096_pdl,full_8192,full_2048_pdl,full_1024 --out gpurun_out/perf_uniform.json 2>&1 | cut -c1-330
