#!/bin/bash
# round 2, run 52 (2 GPUs): bench.py --gpus 2 on the final tree (failure-tolerant e2e leg at N > 1)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_52_bench_n2.json 2> gpurun_out/r2_52_bench_n2.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2_52_bench_n2.json; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r2_52_bench_n2.err | tail -5 | cut -c1-300
