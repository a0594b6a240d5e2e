#!/bin/bash
# round 2, run 50 (1 GPU): the whole GPU suite, smoke(), and both bench arms on the final tree
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_50_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_50_pytest.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_50_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_50_smoke.log | cut -c1-300
timeout 600 python bench.py > gpurun_out/r2_50_bench.json 2> gpurun_out/r2_50_bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r2_50_bench.json
timeout 600 python bench.py --impl reference > gpurun_out/r2_50_bench_reference.json 2> gpurun_out/r2_50_bench_reference.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/r2_50_bench_reference.json
