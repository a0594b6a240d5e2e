#!/bin/bash
mkdir -p gpurun_out
for mb in 0 40 64 0 40; do echo "== QG_L2_PERSIST_MB=$mb"; QG_L2_PERSIST_MB=$mb timeout 300 python tools/gpu_perf.py --only full_4096_pdl,full_8192,full_2048_pdl --out gpurun_out/perf_l2p$mb.json 2>&1 | cut -c1-260; done
QG_L2_PERSIST_MB=40 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
