#!/bin/bash
# round 2, run 23 (1 GPU): outlier path after the LDS fix -- parity, sweep, per-kernel durations; bench with the warm-up order fixed
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_outlier.py -m gpu -x -q -p no:cacheprovider > gpurun_out/r2_23_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_23_pytest.log | cut -c1-200
python - <<'PY' > gpurun_out/r2_23_outlier_sweep.json 2> gpurun_out/r2_23_outlier_sweep.err
import sys, json
sys.path.insert(0, "tools"); sys.path.insert(0, ".")
import bench_configs as bc
print(json.dumps(bc.outlier_sweep_case()))
PY
cat gpurun_out/r2_23_outlier_sweep.json | cut -c1-600; tail -2 gpurun_out/r2_23_outlier_sweep.err
python tools/prof_outlier.py > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_23_outlier_launches.csv python tools/prof_outlier.py > gpurun_out/r2_23_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/r2_23_outlier_launches.csv")) if len(r)>10 and r[0].isdigit()]
half=len(rows)//2
for r in rows[half:]:
    print(r[4][:70].replace("void qg::<unnamed>::",""), r[-1])
PY
for i in 1 2 3; do timeout 300 python bench.py --sustained-seconds 0 --no-cpu-baseline > gpurun_out/r2_23_bench_$i.json 2>/dev/null; python -c "
import json
b=json.loads([l for l in open('gpurun_out/r2_23_bench_$i.json') if l.startswith('{')][-1]); print('bench', $i, round(b['ms_per_step']*1e3,2), 'us', round(b['value'],1), 'TOPS; fp16', round(b['library_context']['cublas_fp16_ms']*1e3,1))"; done
for i in 1 2; do QG_LIB=libqgemm_ul4.so timeout 300 python bench.py --sustained-seconds 0 --no-cpu-baseline > gpurun_out/r2_23_bench_ul4_$i.json 2>/dev/null; python -c "
import json
b=json.loads([l for l in open('gpurun_out/r2_23_bench_ul4_$i.json') if l.startswith('{')][-1]); print('bench ul4', $i, round(b['ms_per_step']*1e3,2), 'us', round(b['value'],1), 'TOPS; fused stage', round(b['stages']['quant_rows_and_cols']['ms']*1e3,2))"; done
python -c "
import json
b=json.loads([l for l in open('gpurun_out/r2_23_bench_1.json') if l.startswith('{')][-1]); print('ul8 fused stage', round(b['stages']['quant_rows_and_cols']['ms']*1e3,2))"
