#!/bin/bash
# round 2, run 21 (8 GPUs): Megatron pairing, exchange tail pipelined over 1 / 2 / 4 / 8 row blocks at P = 8
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29628 tools/bench_megatron.py --only-fused > gpurun_out/r2_21_megatron8.log 2>&1; echo "megatron 8 rc=$?"; tail -1 gpurun_out/r2_21_megatron8.log | cut -c1-1800
