#!/bin/bash
# round 2, run 18 (1 GPU): fused quantizers -- resident CTAs per SM (4 / 5 / 6) x pass-2 share, whole op at 4096^3
mkdir -p gpurun_out
for lib in libqgemm_f4.so libqgemm_f5.so libqgemm.so; do for pct in 40 50 60; do
QG_LIB=$lib QG_FUSED_COLS_PCT=$pct timeout 300 python bench.py --steps 30 --warmup 5 --sustained-seconds 0 > gpurun_out/r2_18_bench_${lib}_$pct.json 2>/dev/null; echo "$lib $pct rc=$?"
done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_18_bench_*.json")):
    try:
        b=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("bench_")[1], "us/step", round(b["ms_per_step"]*1e3,2), "TOPS", round(b["value"],1), "parity", b["parity_checked"], "fp16 us", round(b["library_context"]["cublas_fp16_ms"]*1e3,1))
    except Exception as e: print(f, "ERR", e)
PY
