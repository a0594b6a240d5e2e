#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
NCCL_DEBUG=INFO NCCL_DEBUG_FILE=gpurun_out/nccl_%h_%p.log timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/n2_diag.py > gpurun_out/n2_diag.log 2>&1; echo "diag rc=$?"; grep '^{' gpurun_out/n2_diag.log; tail -5 gpurun_out/n2_diag.log | cut -c1-300
grep -hE "via|NVLS|P2P|SHM|channels|Connected" gpurun_out/nccl_*.log | cut -c1-200 | sort | uniq -c | sort -rn | head -12
for v in a b; do
  if [ $v = b ]; then export QG_BENCH_NO_CLOCKS=1; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 50 --warmup 10 > gpurun_out/bench_n2_$v.json 2> gpurun_out/bench_n2_$v.err
  QG_BENCH_EXCHANGE=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 50 --warmup 10 > gpurun_out/bench_n2_nccl_$v.json 2> gpurun_out/bench_n2_nccl_$v.err
done
unset QG_BENCH_NO_CLOCKS
for f in bench_n2_a bench_n2_nccl_a bench_n2_b bench_n2_nccl_b; do python - <<PY
import json
try:
    b=json.loads([l for l in open("gpurun_out/$f.json") if l.startswith("{")][-1])
    print("$f","value",round(b["value"],1),"ms/step",round(b["ms_per_step"]*1e3,1),"gemm us",round(b["roofline"]["ms"]*1e3,1),"host_enqueue_ms",round(b["host_enqueue_ms"],2),b["config"].get("exchange","")[:40])
except Exception as e:
    print("$f ERR",e); print(open("gpurun_out/$f.err").read()[-1200:])
PY
done
