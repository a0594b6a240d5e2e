#!/bin/bash
# final-state evidence: smoke, full GPU suite, bench arms, launch list (ncu, single-pass metrics, caches as the pipeline leaves them)
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"
timeout 300 python bench.py --impl reference-gpu --steps 10 --warmup 3 > gpurun_out/bench_reference_gpu.json 2> gpurun_out/bench_reference_gpu.err; echo "refgpu rc=$?"
timeout 300 python bench.py > gpurun_out/bench_ours_default.json 2> gpurun_out/bench_ours_default.err; echo "bench default rc=$?"
timeout 300 python bench.py --steps 50 --warmup 10 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench rc=$?"
timeout 300 python bench.py --size 8192 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_ours_8192.json 2> gpurun_out/bench_ours_8192.err
python - <<'PY'
import json
for f in ["bench_ours_default","bench_ours","bench_reference","bench_reference_gpu","bench_ours_8192"]:
    try:
        b=json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        print(f, "value", round(b.get("value"),2), "ms", round(b.get("ms_per_step"),4), "e2e", (b.get("e2e") or {}).get("value"), "roof", (b.get("roofline") or {}).get("frac"), "launches", b.get("gpu_launches"), "clk", (b.get("clocks") or {}).get("sm_mhz"))
    except Exception as e:
        print(f, "ERR", e)
PY
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct -c 120 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
echo "ncu rc=$?"
