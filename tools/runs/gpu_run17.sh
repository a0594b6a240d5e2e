#!/bin/bash
# epilogue rework (pipelined TMEM loads, LDS.128 scales, relaxed tmem-empty arrive): parity, then timings
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python tools/gpu_perf.py --only stats_2sm_4096_f32,stats_2sm_4096_s32,stats_2sm_4096_f16,stats_2sm_4096_f32_mnmajor,stats_2sm_8192_f16,stats_2sm_8192_f32,stats_2sm_2048_f32,stats_2sm_1024_f32,stats_1sm_4096_f32,full_4096_pdl,full_8192,full_2048_pdl,full_1024 --out gpurun_out/perf_epi.json 2>&1 | cut -c1-400
python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench rc=$?"; cat gpurun_out/bench_ours.json | cut -c1-1500
