#!/bin/bash
# early PDL trigger in every kernel: parity, then same-box A/B against the build without it
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
EXPS=full_4096_pdl,full_8192,full_2048_pdl,full_1024,stats_2sm_4096_f32,quant_4096
for rep in 1 2; do
timeout 300 python tools/gpu_perf.py --only $EXPS --out gpurun_out/perf_early$rep.json 2>&1 | cut -c1-330
QG_LIB=libqgemm_noearly.so timeout 300 python tools/gpu_perf.py --only $EXPS --out gpurun_out/perf_noearly$rep.json 2>&1 | cut -c1-330
done
timeout 300 python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_ours.json
QG_LIB=libqgemm_noearly.so timeout 300 python bench.py > gpurun_out/bench_noearly.json 2> gpurun_out/bench_noearly.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_noearly.json
