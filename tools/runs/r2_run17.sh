#!/bin/bash
# round 2, run 17 (1 GPU): the op's quantizers fused (column pass 2 side by side with the row quantizer) -- parity, whole-op timing A/B, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference.py tests/test_gpu_dropin.py -m gpu -x -q -p no:cacheprovider > gpurun_out/r2_17_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_17_pytest.log | cut -c1-300
for pct in 50 35 65; do
QG_FUSED_COLS_PCT=$pct timeout 300 python bench.py --steps 30 --warmup 5 --sustained-seconds 0 > gpurun_out/r2_17_bench_fused$pct.json 2> gpurun_out/r2_17_bench_fused$pct.err; echo "bench fused $pct rc=$?"
done
QG_NO_FUSED_QUANT=1 timeout 300 python bench.py --steps 30 --warmup 5 --sustained-seconds 0 > gpurun_out/r2_17_bench_unfused.json 2> gpurun_out/r2_17_bench_unfused.err; echo "bench unfused rc=$?"
timeout 300 python bench.py -m 8192 -n 8192 -k 8192 --steps 10 --warmup 3 --sustained-seconds 0 > gpurun_out/r2_17_bench_8192_fused.json 2>/dev/null; echo "rc=$?"
QG_NO_FUSED_QUANT=1 timeout 300 python bench.py -m 8192 -n 8192 -k 8192 --steps 10 --warmup 3 --sustained-seconds 0 > gpurun_out/r2_17_bench_8192_unfused.json 2>/dev/null; echo "rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_17_bench_*.json")):
    try:
        b=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("bench_")[1], "us/step", round(b["ms_per_step"]*1e3,2), "TOPS", round(b["value"],1), "parity", b["parity_checked"], "fp16 us", round(b["library_context"]["cublas_fp16_ms"]*1e3,1), "launches", b["gpu_launches"])
    except Exception as e: print(f, "ERR", e)
PY
