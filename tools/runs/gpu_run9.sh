#!/bin/bash
# 2-GPU box: column-parallel tests, multi-GPU bench, plus the single-GPU suite and bench.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_multi.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_multi.log
tail -5 gpurun_out/pytest_multi.log | cut -c1-300
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_gpu_multi.py > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log | cut -c1-300
python bench.py --steps 50 --warmup 10 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 50 --warmup 10 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "n2 rc=$?"
cut -c1-500 gpurun_out/bench_n1.json; cut -c1-500 gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
timeout 600 python tools/gpu_perf.py --only quant_4096,full_4096_pdl,full_2048_pdl,full_8192 > gpurun_out/perf.log 2>&1; cut -c1-400 gpurun_out/perf.log
