#!/bin/bash
mkdir -p gpurun_out
python tools/prof_quant.py > gpurun_out/plain_q.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'quant_cols|absmax_cols|quant_rows' -s 6 -c 3 -o gpurun_out/prof_quant -f python tools/prof_quant.py > gpurun_out/ncu_q.log 2>&1
tail -3 gpurun_out/ncu_q.log
