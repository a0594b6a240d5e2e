#!/bin/bash
# round 2, run 6 (1 GPU): full GPU suite on the tree with the elementwise tail, C++ stack and the cross-layer quantization; bench both arms
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r2_06_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_06_pytest.log | cut -c1-300
timeout 300 python bench.py --impl reference > gpurun_out/r2_06_bench_ref.json 2> gpurun_out/r2_06_bench_ref.err; echo "bench ref rc=$?"
timeout 600 python bench.py > gpurun_out/r2_06_bench.json 2> gpurun_out/r2_06_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r2_06_bench.err
python - <<'PY'
import json
try:
    b=json.loads([l for l in open("gpurun_out/r2_06_bench.json") if l.startswith("{")][-1])
    print("value",round(b["value"],1),"us/step",round(b["ms_per_step"]*1e3,2),"gemm us",round(b["roofline"]["ms"]*1e3,2),"frac",round(b["roofline"]["frac"],3),
          "parity",b["parity_checked"],"e2e",round(b["e2e"]["value"],1),"e2e frac",round(b["e2e"]["roofline"]["frac"],3),"cpu",round(b["cpu_baseline"]["value"],2),
          "sust",round(b["sustained"]["ms_per_step"]*1e3,1), round(b["sustained"]["gemm"]["frac"],3), "fp16 us", round(b["library_context"]["cublas_fp16_ms"]*1e3,1))
    print({k:round(v["ms"]*1e3,2) for k,v in b["stages"].items()})
except Exception as e: print("ERR",e)
PY
