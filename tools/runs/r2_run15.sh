#!/bin/bash
# round 2, run 15 (1 GPU): column quantizer as wavefront launches (pass 2 of slab i-1 rides with pass 1 of slab i) -- timing and DRAM bytes
mkdir -p gpurun_out
timeout 600 python tools/gpu_perf.py --only r2_cwave --out gpurun_out/r2_15_perf.json > gpurun_out/r2_15_perf.log 2>&1; echo "perf rc=$?"
python - <<'PY'
import json
p=json.load(open("gpurun_out/r2_15_perf.json"))
for k,v in p.items(): print(k, {a:(round(b,1) if isinstance(b,float) else b) for a,b in v.items() if a in ("cols_us","digest","graph_replay_ok","error")})
PY
cat > /tmp/prof_cols.py <<'PY'
import sys, importlib, torch
sys.path.insert(0, ".")
qg = importlib.import_module("quantized-gemm-for-transformer-inference_b200")
K = N = 4096
Ws = [torch.rand((K, N), device="cuda") * 2 - 1 for _ in range(3)]
Wq = torch.empty((K, N), dtype=torch.int8, device="cuda"); Cw = torch.empty(N, device="cuda")
for i in range(3): qg.absmax_quant_cols(Ws[i % 3], 127.0, 0, Wq, Cw)
torch.cuda.synchronize(); print("ok")
PY
for w in 2 4; do
QG_COLS_PIPE=0 QG_COLS_WAVE=$w ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none --cache-control none -k regex:cols_wave -c 15 --csv --log-file gpurun_out/r2_15_ncu_wave$w.csv python /tmp/prof_cols.py > gpurun_out/r2_15_ncu.log 2>&1; echo "ncu rc=$?"
grep "gpu__time\|dram__bytes_read\|hit_rate" gpurun_out/r2_15_ncu_wave$w.csv | awk -F'","' '{print $(NF-2), $NF}' | tail -$((3*(w+1)))
done
