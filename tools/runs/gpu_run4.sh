#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 2400 python tools/gpu_perf.py > gpurun_out/perf.log 2>&1
python bench.py --steps 50 --warmup 10 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
cut -c1-400 gpurun_out/bench_ours.json
