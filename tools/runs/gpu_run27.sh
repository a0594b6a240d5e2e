#!/bin/bash
mkdir -p gpurun_out
for lib in libqgemm.so libqgemm_keep.so libqgemm.so libqgemm_keep.so; do echo "== $lib"; QG_LIB=$lib timeout 300 python tools/gpu_perf.py --only quant_4096,full_4096_pdl,full_8192,full_2048_pdl --out gpurun_out/perf_$lib.json 2>&1 | grep -v "twopass\|f16" | cut -c1-330; done
QG_LIB=libqgemm_keep.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
