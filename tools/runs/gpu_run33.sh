#!/bin/bash
mkdir -p gpurun_out
python tools/prof_stack.py > gpurun_out/plain_s.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'add_layernorm|softmax_warp|mm_f32' -s 6 -c 6 -o gpurun_out/prof_stack -f python tools/prof_stack.py > gpurun_out/ncu_s.log 2>&1
tail -2 gpurun_out/ncu_s.log
