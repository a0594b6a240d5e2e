#!/bin/bash
# round 2, run 19 (8 GPUs): multi-GPU parity (column-parallel unicast / multicast, Megatron), scaling of the bench with the multicast
# exchange and with the unicast one, the OPT-66B FFN as a Megatron pair and as two column-parallel layers at P = 8 and 4
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8 > gpurun_out/r2_19_smi.txt 2>&1
nvidia-smi topo -m > gpurun_out/r2_19_topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_megatron.py -m gpu -x -q -p no:cacheprovider -k "multi or fused_exchange or both_quantizers" > gpurun_out/r2_19_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_19_pytest.log | cut -c1-300
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 8 4 2; do
timeout 600 $TR --nproc-per-node $n --master-port 2960$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2_19_bench_n${n}_mc.json 2> gpurun_out/r2_19_bench_n${n}_mc.err; echo "bench n=$n mc rc=$?"
done
QG_NO_MULTICAST=1 timeout 600 $TR --nproc-per-node 8 --master-port 29618 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_19_bench_n8_unicast.json 2> gpurun_out/r2_19_bench_n8_unicast.err; echo "bench n=8 unicast rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_19_bench_n1.json 2> gpurun_out/r2_19_bench_n1.err; echo "bench n=1 rc=$?"
timeout 900 $TR --nproc-per-node 8 --master-port 29628 tools/bench_megatron.py > gpurun_out/r2_19_megatron8.log 2>&1; echo "megatron 8 rc=$?"; tail -1 gpurun_out/r2_19_megatron8.log | cut -c1-1400
timeout 900 $TR --nproc-per-node 4 --master-port 29624 tools/bench_megatron.py > gpurun_out/r2_19_megatron4.log 2>&1; echo "megatron 4 rc=$?"; tail -1 gpurun_out/r2_19_megatron4.log | cut -c1-1400
timeout 900 $TR --nproc-per-node 8 --master-port 29638 tools/bench_colpar.py > gpurun_out/r2_19_colpar8.log 2>&1; echo "colpar 8 rc=$?"; tail -2 gpurun_out/r2_19_colpar8.log | cut -c1-700
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_19_bench_n*.json")):
    try:
        b=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("bench_")[1], "n", b["n_gpus"], "us/step", round(b["ms_per_step"]*1e3,1), "TOPS", round(b["value"],1), "parity", b["parity_checked"], "gemm us", round(b["roofline"]["ms"]*1e3,1), b.get("exchange_used","")[:60], (b.get("exchange_roofline") or {}).get("frac"))
    except Exception as e: print(f, "ERR", e)
PY
