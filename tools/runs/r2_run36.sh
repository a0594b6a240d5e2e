#!/bin/bash
# round 2, run 36 (1 GPU): attention core in one kernel, 64-row tiles at two CTAs per SM -- parity (attention, transformer, drop-in tests), block timing; ncu source-level capture of ADD & NORM
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_transformer.py tests/test_gpu_dropin.py -m gpu -x -q -p no:cacheprovider > gpurun_out/r2_36_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_36_pytest.log | cut -c1-300
python - <<'PY' > gpurun_out/r2_36_stack.json 2> gpurun_out/r2_36_stack.err
import sys, json
sys.path.insert(0, "tools"); sys.path.insert(0, ".")
import bench_configs as bc
print(json.dumps(bc.attention_case()))
print(json.dumps(bc.transformer_case()))
PY
cat gpurun_out/r2_36_stack.json | cut -c1-400; tail -2 gpurun_out/r2_36_stack.err
python tools/prof_block.py > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:attention_core -c 4 --log-file gpurun_out/r2_36_attn_launches.csv python tools/prof_block.py > gpurun_out/r2_36_ncu.log 2>&1; grep attention_core gpurun_out/r2_36_attn_launches.csv | awk -F'","' '{print "attention_core_kernel ns", $NF}' | tail -3
ncu --set full --clock-control none --import-source on -k regex:add_layernorm_cta_rows -s 4 -c 1 -f -o gpurun_out/r2_36_addnorm python tools/prof_block.py > gpurun_out/r2_36_ncu2.log 2>&1; echo "ncu addnorm rc=$?"
