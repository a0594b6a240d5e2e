#!/bin/bash
# round 2, run 30 (2 GPUs): every GPU test on the final tree (single- and multi-GPU)
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_30_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2_30_pytest.log | cut -c1-300
