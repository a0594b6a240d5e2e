#!/bin/bash
# round 2, run 54 (1 GPU): smoke() and the Megatron / transformer / parity files on the final library
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_54_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2_54_smoke.log | cut -c1-200
timeout 400 python -m pytest tests/test_gpu_megatron.py tests/test_gpu_transformer.py tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider > gpurun_out/r2_54_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_54_pytest.log | cut -c1-200
