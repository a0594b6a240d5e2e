#!/bin/bash
mkdir -p gpurun_out/golden
python tests/golden/make_ref_fixtures.py gpurun_out/golden > gpurun_out/fixtures.log 2>&1; echo "fixtures rc=$?"; tail -3 gpurun_out/fixtures.log
cp gpurun_out/golden/ref_enc_*.npz gpurun_out/golden/ref_addnorm_*.npz tests/golden/ 2>/dev/null
timeout 600 python -m pytest tests/test_gpu_transformer.py tests/test_oracle_cpu.py -x -q 2>&1 | tail -15
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
for c in 8 16 4; do QG_HOST_CHUNKS=$c python tools/e2e_probe.py; done
