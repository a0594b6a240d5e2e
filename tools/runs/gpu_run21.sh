#!/bin/bash
# pipelined host entry point: parity + e2e; final-state evidence for profiles/ (bench arms, one ncu --set full capture)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 300 python bench.py --steps 50 --warmup 10 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"
timeout 300 python bench.py --impl reference-gpu --steps 10 --warmup 3 > gpurun_out/bench_reference_gpu.json 2> gpurun_out/bench_reference_gpu.err; echo "refgpu rc=$?"
timeout 300 python bench.py --size 8192 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_ours_8192.json 2> gpurun_out/bench_ours_8192.err
python - <<'PY'
import json
for f in ["bench_ours","bench_reference","bench_reference_gpu","bench_ours_8192"]:
    try:
        b=json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        print(f, "value", b.get("value"), "ms", b.get("ms_per_step"), "e2e", b.get("e2e"), "roof", (b.get("roofline") or {}).get("frac"), (b.get("roofline") or {}).get("ms"))
    except Exception as e:
        print(f, "ERR", e)
PY
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'gemm_i8_tc|quant_rows_kernel|quant_cols_kernel|absmax_cols_partial' -s 8 -c 8 \
    -o gpurun_out/prof_r1_final -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
