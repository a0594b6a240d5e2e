#!/bin/bash
# round 2, run 20 (2 GPUs): Megatron pairing with the exchange tail pipelined over row blocks -- parity, timing at P = 2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_megatron.py tests/test_gpu_multi.py -m gpu -x -q -p no:cacheprovider > gpurun_out/r2_20_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_20_pytest.log | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29574 tools/bench_megatron.py > gpurun_out/r2_20_megatron.log 2>&1; echo "full rc=$?"; tail -1 gpurun_out/r2_20_megatron.log | cut -c1-1800
