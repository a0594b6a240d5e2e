#!/bin/bash
# 8-GPU box: scaling lines of bench.py (fused symmetric-memory exchange, and NCCL all-gather for comparison)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
run() { # name nproc extra-env
  env $3 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus $2 --steps 30 --warmup 5 > gpurun_out/$1.json 2> gpurun_out/$1.err
  echo "$1 rc=$?"
  python - <<PY
import json
try:
    b=json.loads([l for l in open("gpurun_out/$1.json") if l.startswith("{")][-1])
    print("$1","value",round(b["value"],1),"us/step",round(b["ms_per_step"]*1e3,1),"gemm us",round(b["roofline"]["ms"]*1e3,1),b["config"].get("exchange","")[:60], b.get("clocks"))
except Exception as e:
    print("$1 ERR",e); print(open("gpurun_out/$1.err").read()[-1500:])
PY
}
run bench_n8 8 QG_X=1
run bench_n4 4 QG_X=1
run bench_n2 2 QG_X=1
run bench_n8_nccl 8 QG_BENCH_EXCHANGE=nccl
run bench_n4_nccl 4 QG_BENCH_EXCHANGE=nccl
timeout 200 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "n1 rc=$?"; cut -c1-300 gpurun_out/bench_n1.json
