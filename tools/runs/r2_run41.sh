#!/bin/bash
# round 2, run 41 (1 GPU): config-3 stack as one CUDA graph vs call by call
mkdir -p gpurun_out
timeout 300 python tools/prof_stack_graph.py > gpurun_out/r2_41_stack_graph.json 2> gpurun_out/r2_41_stack_graph.err; echo "rc=$?"; cat gpurun_out/r2_41_stack_graph.json; tail -5 gpurun_out/r2_41_stack_graph.err | cut -c1-400
