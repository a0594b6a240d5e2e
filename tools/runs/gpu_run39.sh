#!/bin/bash
# BASELINE config 4 on one 8-GPU box: OPT-66B FFN column-parallel at P = 8, 4, 2
mkdir -p gpurun_out
for P in 8 4 2; do
  timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1 --master-port $((29570 + P)) tools/bench_colpar.py > gpurun_out/colpar_$P.log 2>&1; echo "P=$P rc=$?"
  grep '^{' gpurun_out/colpar_$P.log | cut -c1-400; tail -3 gpurun_out/colpar_$P.log | grep -v '^{' | cut -c1-200
done
