#!/bin/bash
# GPU run 2: smoke, C++ drop-in driver, full GPU test suite, bench (both arms), then ncu.
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
./tests/cpp/test_quantize_dropin > gpurun_out/dropin_cpp.log 2>&1; echo "rc=$?" >> gpurun_out/dropin_cpp.log
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python bench.py --impl reference-gpu --steps 10 --warmup 3 > gpurun_out/bench_reference_gpu.json 2> gpurun_out/bench_reference_gpu.err
python bench.py --steps 50 --warmup 10 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
QG_GEMM_VARIANT=3 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/bench_ours_2sm.json 2> gpurun_out/bench_ours_2sm.err
python bench.py --size 8192 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_ours_8192.json 2> gpurun_out/bench_ours_8192.err
QG_GEMM_VARIANT=3 python bench.py --size 8192 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_ours_8192_2sm.json 2> gpurun_out/bench_ours_8192_2sm.err
# ncu: launch list, then one full capture of the hot kernels (same command line as the plain run)
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'gemm_i8_tc|quant_rows_kernel|quant_cols_kernel|absmax_cols_partial' -s 12 -c 8 \
    -o gpurun_out/prof_r1 -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log gpurun_out/dropin_cpp.log | tail -8
cat gpurun_out/bench_ours.json | cut -c1-1500
