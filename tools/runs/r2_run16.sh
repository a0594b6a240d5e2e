#!/bin/bash
# round 2, run 16 (1 GPU): L2 reuse between the two column-quantizer passes, timed without a profiler; compute-sanitizer on the smoke shape
mkdir -p gpurun_out
timeout 600 python tools/l2_reuse_probe.py > gpurun_out/r2_16_l2_reuse.log 2>&1; echo "probe rc=$?"; cat gpurun_out/r2_16_l2_reuse.log | tail -8
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/r2_16_memcheck_smoke.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r2_16_memcheck_smoke.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/r2_16_racecheck_smoke.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r2_16_racecheck_smoke.log
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/r2_16_synccheck_smoke.log 2>&1; echo "synccheck rc=$?"; tail -4 gpurun_out/r2_16_synccheck_smoke.log
