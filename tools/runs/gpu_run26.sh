#!/bin/bash
# 2-GPU box: multi-GPU tests + bench at N=2 (fused exchange and NCCL all-gather) + N=1 on the same box
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_multi.log 2>&1; echo "multi rc=$?"; tail -2 gpurun_out/pytest_multi.log | cut -c1-200
timeout 300 python bench.py --steps 50 --warmup 10 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "n1 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 50 --warmup 10 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "n2 rc=$?"
QG_BENCH_EXCHANGE=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 50 --warmup 10 > gpurun_out/bench_n2_nccl.json 2> gpurun_out/bench_n2_nccl.err; echo "n2 nccl rc=$?"
for f in bench_n1 bench_n2 bench_n2_nccl; do python - <<PY
import json
try:
    b=json.loads([l for l in open("gpurun_out/$f.json") if l.startswith("{")][-1])
    print("$f","value",round(b["value"],1),"us/step",round(b["ms_per_step"]*1e3,1),"gemm us",round(b["roofline"]["ms"]*1e3,1),"host_enqueue_ms",round(b["host_enqueue_ms"],2),b["config"].get("exchange","")[:40], b["clocks"]["sm_mhz"], b["clocks"]["samples"], "e2e", (b.get("e2e") or {}).get("value"))
except Exception as e:
    print("$f ERR",e); print(open("gpurun_out/$f.err").read()[-1200:])
PY
done
