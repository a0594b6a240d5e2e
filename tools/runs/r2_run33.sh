#!/bin/bash
# round 2, run 33 (8 GPUs): Megatron pairing at P = 8 -- row blocks x gather engine (peer stores from the reduce kernel / copy engines)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29628 tools/bench_megatron.py --only-fused > gpurun_out/r2_33_megatron8.log 2>&1; echo "megatron 8 rc=$?"; tail -1 gpurun_out/r2_33_megatron8.log | cut -c1-1800
