#!/bin/bash
# round 2, run 29 (2 GPUs): every GPU test on the final tree (single- and multi-GPU), smoke, bench at N = 1 and N = 2
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r2_29_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_29_pytest.log | cut -c1-300
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_29_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2_29_smoke.log
timeout 600 python bench.py > gpurun_out/r2_29_bench_n1.json 2> gpurun_out/r2_29_bench_n1.err; echo "bench n1 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 bench.py --gpus 2 > gpurun_out/r2_29_bench_n2.json 2> gpurun_out/r2_29_bench_n2.err; echo "bench n2 rc=$?"
python tools/prof_outlier.py > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:"outlier_detect|outlier_index" -c 6 --log-file gpurun_out/r2_29_detect_launches.csv python tools/prof_outlier.py > gpurun_out/r2_29_ncu.log 2>&1; echo "ncu rc=$?"; grep "outlier_" gpurun_out/r2_29_detect_launches.csv | awk -F'","' '{print substr($5,1,50), $NF}' | tail -4
python - <<'PY'
import json
for n in ("n1","n2"):
    try:
        b=json.loads([l for l in open(f"gpurun_out/r2_29_bench_{n}.json") if l.startswith("{")][-1])
        print(n, "value",round(b["value"],1),"us/step",round(b["ms_per_step"]*1e3,2),"gemm us",round(b["roofline"]["ms"]*1e3,2),"frac",round(b["roofline"]["frac"],3),"parity",b["parity_checked"], (b.get("exchange_used") or "")[:50])
        if n=="n1": print("   e2e",round(b["e2e"]["value"],1),"e2e frac",round(b["e2e"]["roofline"]["frac"],3),"cpu",round(b["cpu_baseline"]["value"],2),"sust",round(b["sustained"]["ms_per_step"]*1e3,1),"fp16 us", round(b["library_context"]["cublas_fp16_ms"]*1e3,1), {k:(round(v["ms"]*1e3,2) if v else None) for k,v in b["stages"].items()})
    except Exception as e: print(n, "ERR", e)
PY
