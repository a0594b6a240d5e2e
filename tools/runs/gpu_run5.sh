#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference.py -m gpu -q -p no:cacheprovider -x -k "col_quantizer or prepare_weights or quantized_mm or reference or linear or golden" > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 1200 python tools/gpu_perf.py --only quant_,full_4096_pdl,full_8192,full_2048_pdl,full_1024 > gpurun_out/perf.log 2>&1
cat gpurun_out/perf.log | cut -c1-400
