#!/bin/bash
# round 2, run 10 (1 GPU): column pipeline with per-warp work items (no block barriers) -- parity, sweep, DRAM bytes
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference.py -m gpu -x -q -p no:cacheprovider > gpurun_out/r2_10_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_10_pytest.log | cut -c1-300
timeout 1500 python tools/gpu_perf.py --only r2_cpipe,r2_cdbg --out gpurun_out/r2_10_perf.json > gpurun_out/r2_10_perf.log 2>&1; echo "perf rc=$?"
python - <<'PY'
import json
p=json.load(open("gpurun_out/r2_10_perf.json"))
for k,v in p.items(): print(k, {a:(round(b,1) if isinstance(b,float) else b) for a,b in v.items() if a in ("cols_us","digest","graph_replay_ok","error")})
PY
cat > /tmp/prof_cols.py <<'PY'
import sys, importlib, torch
sys.path.insert(0, ".")
qg = importlib.import_module("quantized-gemm-for-transformer-inference_b200")
K = N = 4096
Ws = [torch.rand((K, N), device="cuda") * 2 - 1 for _ in range(3)]
Wq = torch.empty((K, N), dtype=torch.int8, device="cuda"); Cw = torch.empty(N, device="cuda")
for i in range(6): qg.absmax_quant_cols(Ws[i % 3], 127.0, 0, Wq, Cw)
torch.cuda.synchronize(); print("ok")
PY
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:quant_cols -c 6 --csv --log-file gpurun_out/r2_10_ncu_cols.csv python /tmp/prof_cols.py > gpurun_out/r2_10_ncu.log 2>&1; echo "ncu rc=$?"
grep "gpu__time\|dram__bytes_read" gpurun_out/r2_10_ncu_cols.csv | awk -F'","' '{print $(NF-2), $NF}' | head -6
