#!/bin/bash
# GPU run 3: correctness of the reworked quantizers / PDL, then the perf experiments and bench.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
QG_PDL=0 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "quantizer or quantized_mm" > gpurun_out/pytest_nopdl.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_nopdl.log
QG_COLS_TWO_PASS=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "col_quantizer" > gpurun_out/pytest_twopass.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_twopass.log
tail -2 gpurun_out/pytest_nopdl.log gpurun_out/pytest_twopass.log
timeout 2400 python tools/gpu_perf.py > gpurun_out/perf.log 2>&1
python bench.py --steps 50 --warmup 10 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
cat gpurun_out/bench_ours.json | cut -c1-600
