#!/bin/bash
# First GPU bring-up: diagnostics (isolated subprocesses), then the parity suites per kernel variant.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.txt 2>&1
nproc >> gpurun_out/smi.txt; lscpu | grep "Model name" >> gpurun_out/smi.txt
timeout 1500 python tools/gpu_diag.py > gpurun_out/diag.log 2>&1
echo "diag rc=$?" >> gpurun_out/diag.log
timeout 300 python tests/golden/make_ref_fixtures.py gpurun_out/golden > gpurun_out/fixtures.log 2>&1
echo "fixtures rc=$?" >> gpurun_out/fixtures.log
timeout 900 python -m pytest tests -m gpu -q -k "not TC_1SM and not TC_2SM and not kmajor and not large_sampled" -p no:cacheprovider > gpurun_out/pytest_base.log 2>&1
echo "rc=$?" >> gpurun_out/pytest_base.log
timeout 900 python -m pytest tests -m gpu -q -k "TC_1SM or (kmajor and 1)" -p no:cacheprovider > gpurun_out/pytest_tc1.log 2>&1
echo "rc=$?" >> gpurun_out/pytest_tc1.log
timeout 900 python -m pytest tests -m gpu -q -k "TC_2SM or (kmajor and 2)" -p no:cacheprovider > gpurun_out/pytest_tc2.log 2>&1
echo "rc=$?" >> gpurun_out/pytest_tc2.log
tail -3 gpurun_out/pytest_base.log gpurun_out/pytest_tc1.log gpurun_out/pytest_tc2.log
grep -c '"ok": true' gpurun_out/diag.json
