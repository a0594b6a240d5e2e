#!/bin/bash
# round 2, run 51 (2 GPUs): the multi-GPU test files and the two-device test on the final tree
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_multi.py tests/test_gpu_megatron.py tests/test_gpu_parity.py -k "multi or megatron or second_device or Multi or Megatron" -m gpu -q -p no:cacheprovider > gpurun_out/r2_51_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_51_pytest.log | cut -c1-300
