#!/bin/bash
# round 2, run 45 (1 GPU): ncu source-level capture of the attention core inside a block (caches as the pipeline leaves them)
mkdir -p gpurun_out
ncu --set full --clock-control none --cache-control none --import-source on -k regex:attention_core -s 2 -c 1 -f -o gpurun_out/r2_45_attn python tools/prof_block.py > gpurun_out/r2_45_ncu.log 2>&1; echo "ncu rc=$?"
