#!/bin/bash
# round 2, run 31 (2 GPUs): Megatron pairing with SMs reserved for the exchange tail -- parity (multi-GPU tests), timing at P = 2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_megatron.py -m gpu -x -q -p no:cacheprovider > gpurun_out/r2_31_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_31_pytest.log | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29574 tools/bench_megatron.py --only-fused > gpurun_out/r2_31_megatron.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/r2_31_megatron.log | cut -c1-1800
