#!/bin/bash
# 2-GPU box: triple-buffered epilogue staging for peer stores -- parity, then bench N=2 with and without it
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -p no:cacheprovider 2>&1 | tail -2 | cut -c1-200
for v in on off on off; do
  if [ $v = off ]; then export QG_NO_STAGE_MULTIBUF=1; else unset QG_NO_STAGE_MULTIBUF; fi
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 50 --warmup 10 > gpurun_out/bench_n2_$v.json 2> gpurun_out/bench_n2_$v.err
  python - <<PY
import json
try:
    b=json.loads([l for l in open("gpurun_out/bench_n2_$v.json") if l.startswith("{")][-1])
    print("multibuf $v","value",round(b["value"],1),"us/step",round(b["ms_per_step"]*1e3,1),"gemm us",round(b["roofline"]["ms"]*1e3,1))
except Exception as e:
    print("ERR",e); print(open("gpurun_out/bench_n2_$v.err").read()[-800:])
PY
done
