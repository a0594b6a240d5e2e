#!/bin/bash
# round 2, run 1 (1 GPU): full GPU test suite on the new tree, decoder fixtures from the reference's kernels,
# GEMM experiments (last-tile ring staging, epilogue / load ablations), bench both arms
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_01_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r2_01_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_01_pytest.log | cut -c1-300
timeout 300 python tests/golden/make_ref_fixtures.py gpurun_out/golden > gpurun_out/r2_01_fixtures.log 2>&1; echo "fixtures rc=$?"; tail -3 gpurun_out/r2_01_fixtures.log
timeout 900 python tools/gpu_perf.py --only r2_ --out gpurun_out/r2_01_perf.json > gpurun_out/r2_01_perf.log 2>&1; echo "perf rc=$?"
timeout 300 python bench.py --impl reference > gpurun_out/r2_01_bench_ref.json 2> gpurun_out/r2_01_bench_ref.err; echo "bench ref rc=$?"
timeout 600 python bench.py > gpurun_out/r2_01_bench.json 2> gpurun_out/r2_01_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r2_01_bench.err
python - <<'PY'
import json
try:
    b=json.loads([l for l in open("gpurun_out/r2_01_bench.json") if l.startswith("{")][-1])
    print("value",round(b["value"],1),"us/step",round(b["ms_per_step"]*1e3,2),"gemm us",round(b["roofline"]["ms"]*1e3,2),"frac",round(b["roofline"]["frac"],3),
          "parity",b["parity_checked"],"e2e",round(b["e2e"]["value"],1),"e2e frac",round(b["e2e"]["roofline"]["frac"],3),"cpu",round(b["cpu_baseline"]["value"],2),
          "sust",round(b["sustained"]["ms_per_step"]*1e3,1), round(b["sustained"]["gemm"]["frac"],3), "fp16 us", round(b["library_context"]["cublas_fp16_ms"]*1e3,1))
    print({k:round(v["ms"]*1e3,2) for k,v in b["stages"].items()})
except Exception as e: print("ERR",e)
try:
    r=json.loads([l for l in open("gpurun_out/r2_01_bench_ref.json") if l.startswith("{")][-1]); print("ref",round(r["value"],3),"TOPS",round(r["ms_per_step"],1),"ms", r["cpu_baseline"]["cores"])
except Exception as e: print("ERR ref",e)
try:
    p=json.load(open("gpurun_out/r2_01_perf.json"))
    for k,v in p.items(): print(k, {a:(round(b,1) if isinstance(b,float) else b) for a,b in v.items() if a in ("us","us_with_stats","mma_total","mma_wait_full","epi_total","epi_wait_tfull","gemm_us","cols_us","rows_us","total_us","error","first_start_to_last_end_us")})
except Exception as e: print("ERR perf",e)
PY
