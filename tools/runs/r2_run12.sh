#!/bin/bash
# round 2, run 12 (2 GPUs): Megatron pairing (single-device scatter/reduce parity, 2-GPU fused exchange parity, config-5 FFN timing at P = 2), multicast gather parity, outlier side product up to 64 columns
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2_12_smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_megatron.py tests/test_gpu_multi.py tests/test_gpu_outlier.py -m gpu -x -q -p no:cacheprovider > gpurun_out/r2_12_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2_12_pytest.log | cut -c1-400
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29573 tools/bench_megatron.py --small > gpurun_out/r2_12_megatron_small.log 2>&1; echo "small rc=$?"; tail -3 gpurun_out/r2_12_megatron_small.log | cut -c1-1500
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29574 tools/bench_megatron.py > gpurun_out/r2_12_megatron.log 2>&1; echo "full rc=$?"; tail -3 gpurun_out/r2_12_megatron.log | cut -c1-1500
