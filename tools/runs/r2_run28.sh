#!/bin/bash
# round 2, run 28 (1 GPU): GEMM with two CTA pairs per cluster (512 x 256 cluster tiles, B quarter-loads multicast between the pairs) -- parity subset under a short timeout, then timing
mkdir -p gpurun_out
QG_GEMM_NP=2 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider -k "gemm_s8s8s32_bit_exact or gemm_dequant_epilogue or gemm_large_sampled" > gpurun_out/r2_28_pytest_np2.log 2>&1; echo "pytest np2 rc=$?"; tail -6 gpurun_out/r2_28_pytest_np2.log | cut -c1-400
grep -m3 "mbarrier timeout" gpurun_out/r2_28_pytest_np2.log
QG_GEMM_NP=2 timeout 300 python tools/gpu_perf.py --only r2_stats_2sm_4096_f32,r2_stats_2sm_4096_f32_mn,r2_stats_2sm_8192_f32,r2_inop_4096 --out gpurun_out/r2_28_perf_np2.json > gpurun_out/r2_28_perf_np2.log 2>&1; echo "perf np2 rc=$?"
timeout 300 python tools/gpu_perf.py --only r2_stats_2sm_4096_f32,r2_stats_2sm_4096_f32_mn,r2_stats_2sm_8192_f32,r2_inop_4096 --out gpurun_out/r2_28_perf_np1.json > gpurun_out/r2_28_perf_np1.log 2>&1; echo "perf np1 rc=$?"
python - <<'PY'
import json
for n in ("np2","np1"):
    try:
        p=json.load(open(f"gpurun_out/r2_28_perf_{n}.json"))
        for k,v in p.items(): print(n, k, {a:(round(b,1) if isinstance(b,float) else b) for a,b in v.items() if a in ("us","us_with_stats","mma_total","mma_wait_full","gemm_us","total_us","error","first_start_to_last_end_us")})
    except Exception as e: print(n, "ERR", e)
PY
QG_GEMM_NP=2 timeout 300 python bench.py --sustained-seconds 0 --no-cpu-baseline > gpurun_out/r2_28_bench_np2.json 2>/dev/null; python -c "
import json
b=json.loads([l for l in open('gpurun_out/r2_28_bench_np2.json') if l.startswith('{')][-1]); print('bench np2', round(b['ms_per_step']*1e3,2), 'us', round(b['value'],1), 'TOPS parity', b['parity_checked'], 'gemm', round(b['roofline']['ms']*1e3,2))"
timeout 300 python bench.py --sustained-seconds 0 --no-cpu-baseline > gpurun_out/r2_28_bench_np1.json 2>/dev/null; python -c "
import json
b=json.loads([l for l in open('gpurun_out/r2_28_bench_np1.json') if l.startswith('{')][-1]); print('bench np1', round(b['ms_per_step']*1e3,2), 'us', round(b['value'],1), 'TOPS parity', b['parity_checked'], 'gemm', round(b['roofline']['ms']*1e3,2))"
