#!/bin/bash
# round 2, run 22 (1 GPU): full GPU suite on the final tree, bench both arms, BASELINE configs 0-5 (+ outlier sweep, FFN chain), ncu launch list and --set full captures
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r2_22_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_22_pytest.log | cut -c1-300
timeout 300 python bench.py --impl reference > gpurun_out/r2_22_bench_ref.json 2> gpurun_out/r2_22_bench_ref.err; echo "bench ref rc=$?"
timeout 600 python bench.py > gpurun_out/r2_22_bench.json 2> gpurun_out/r2_22_bench.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r2_22_bench.err
timeout 1200 python tools/bench_configs.py > gpurun_out/r2_22_configs.log 2>&1; echo "configs rc=$?"; tail -8 gpurun_out/r2_22_configs.log | cut -c1-600
cp gpurun_out/configs.json gpurun_out/r2_22_configs.json 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none -c 400 --csv --log-file gpurun_out/r2_22_launches.csv python bench.py --steps 2 --warmup 3 --sustained-seconds 0 --no-cpu-baseline > gpurun_out/r2_22_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"quant_cols2_rows|gemm_i8_tc|absmax_cols_partial" -s 9 -c 3 -f -o gpurun_out/r2_22_op_full python bench.py --steps 2 --warmup 3 --sustained-seconds 0 --no-cpu-baseline > gpurun_out/r2_22_ncu_full.log 2>&1; echo "ncu full rc=$?"
python - <<'PY'
import json
b=json.loads([l for l in open("gpurun_out/r2_22_bench.json") if l.startswith("{")][-1])
print("value",round(b["value"],1),"us/step",round(b["ms_per_step"]*1e3,2),"gemm us",round(b["roofline"]["ms"]*1e3,2),"frac",round(b["roofline"]["frac"],3),
      "parity",b["parity_checked"],"e2e",round(b["e2e"]["value"],1),"e2e frac",round(b["e2e"]["roofline"]["frac"],3),"cpu",round(b["cpu_baseline"]["value"],2),
      "sust",round(b["sustained"]["ms_per_step"]*1e3,1), round(b["sustained"]["gemm"]["frac"],3), "fp16 us", round(b["library_context"]["cublas_fp16_ms"]*1e3,1))
print({k:(round(v["ms"]*1e3,2) if v else None) for k,v in b["stages"].items()})
print(b["linear_prepared_weights"])
PY
