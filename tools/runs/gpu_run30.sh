#!/bin/bash
mkdir -p gpurun_out
for pf in 0 4 8 16 0 8; do echo "== QG_GEMM_PREFETCH=$pf"; QG_GEMM_PREFETCH=$pf timeout 300 python tools/gpu_perf.py --only full_4096_pdl,full_8192,stats_2sm_4096_f32 --out gpurun_out/perf_pf$pf.json 2>&1 | grep -v "nosplit\|mnmajor" | cut -c1-260; done
QG_GEMM_PREFETCH=8 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
