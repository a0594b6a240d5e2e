#!/bin/bash
# round 2, run 9 (1 GPU): where the one-launch column pipeline loses its time -- readers alone, writers alone, ncu --set full
mkdir -p gpurun_out
timeout 600 python tools/gpu_perf.py --only r2_cdbg --out gpurun_out/r2_09_perf.json > gpurun_out/r2_09_perf.log 2>&1; echo "perf rc=$?"
python - <<'PY'
import json
p=json.load(open("gpurun_out/r2_09_perf.json"))
for k,v in p.items(): print(k, {a:(round(b,1) if isinstance(b,float) else b) for a,b in v.items() if a in ("cols_us","cols_gbs","error")})
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
QG_COLS_PIPE=8,64,5,8,8 ncu --set full --clock-control none --import-source on -k regex:quant_cols_pipe -s 3 -c 1 -f -o gpurun_out/r2_09_cols_pipe python /tmp/prof_cols.py > gpurun_out/r2_09_ncu.log 2>&1; echo "ncu rc=$?"
