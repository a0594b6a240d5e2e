#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
python -c "import torch;print('p2p01',torch.cuda.can_device_access_peer(0,1))" >> gpurun_out/topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -p no:cacheprovider -x > gpurun_out/pytest_multi.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_multi.log
tail -30 gpurun_out/pytest_multi.log | cut -c1-300
python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 50 --warmup 10 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "n2 rc=$?"
QG_BENCH_EXCHANGE=nccl NCCL_DEBUG_FILE=gpurun_out/nccl_%h_%p.log timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 50 --warmup 10 > gpurun_out/bench_n2_nccl.json 2> gpurun_out/bench_n2_nccl.err
cat gpurun_out/topo.txt
for f in bench_n1 bench_n2 bench_n2_nccl; do python - <<PY
import json
try:
    b=json.loads([l for l in open("gpurun_out/$f.json") if l.startswith("{")][-1])
    print("$f","value",round(b["value"],1),"ms/step",round(b["ms_per_step"]*1e3,1),"gemm us",round(b["roofline"]["ms"]*1e3,1),b["config"].get("exchange"),b["gpu_launches"])
except Exception as e:
    print("$f ERR",e); print(open("gpurun_out/$f.err").read()[-1500:])
PY
done
