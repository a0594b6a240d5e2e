#!/bin/bash
# round 2, run 27 (1 GPU): masked row quantizer -- mask bits fetched once per thread, register bound for 4 CTAs per SM: parity, sweep, per-kernel durations
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_outlier.py tests/test_gpu_parity.py tests/test_gpu_transformer.py -m gpu -x -q -p no:cacheprovider  > gpurun_out/r2_27_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_27_pytest.log | cut -c1-200
python - <<'PY' > gpurun_out/r2_27_outlier_sweep.json 2> gpurun_out/r2_27_outlier_sweep.err
import sys, json
sys.path.insert(0, "tools"); sys.path.insert(0, ".")
import bench_configs as bc
print(json.dumps(bc.outlier_sweep_case()))
for name, K, N in (("opt6.7b_qkv", 4096, 12288), ("opt6.7b_fc1", 4096, 16384), ("opt6.7b_fc2", 16384, 4096)):
    print(json.dumps(bc.linear_case(name, 16384, K, N, bc.torch.float16, outliers=6)))
PY
cat gpurun_out/r2_27_outlier_sweep.json | cut -c1-700; tail -2 gpurun_out/r2_27_outlier_sweep.err
python tools/prof_outlier.py > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_27_outlier_launches.csv python tools/prof_outlier.py > gpurun_out/r2_27_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/r2_27_outlier_launches.csv")) if len(r)>10 and r[0].isdigit()]
half=len(rows)//2
for r in rows[half:]:
    if "qg::" in r[4]: print(r[4][:70].replace("void qg::<unnamed>::",""), r[-1])
PY
