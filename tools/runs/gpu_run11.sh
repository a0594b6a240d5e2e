#!/bin/bash
mkdir -p gpurun_out
python tools/p2p_check.py > gpurun_out/p2p.txt 2>&1; cat gpurun_out/p2p.txt | tail -25
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_multi.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_multi.log
tail -6 gpurun_out/pytest_multi.log | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -x -k "col_quantizer or prepare_weights or quantized_mm or linear or golden or prepared" > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 600 python tools/gpu_perf.py --only quant_4096,quant_8192,full_4096_pdl,full_2048_pdl > gpurun_out/perf.log 2>&1; cut -c1-420 gpurun_out/perf.log
QG_PERCALL_KMAJOR=0 timeout 600 python tools/gpu_perf.py --only full_4096_pdl,full_2048_pdl --out gpurun_out/perf_mn.json > gpurun_out/perf_mn.log 2>&1; cut -c1-420 gpurun_out/perf_mn.log
