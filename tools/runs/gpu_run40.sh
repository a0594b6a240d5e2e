#!/bin/bash
# split-K: parity (whole GPU suite; a forced split factor over the GEMM tests as well), then the tile-starved shapes
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
QG_SPLIT_K=3 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gemm or linear or quantized_mm" 2>&1 | tail -2
python tools/shape_probe.py; QG_SPLIT_K=1 python tools/shape_probe.py
