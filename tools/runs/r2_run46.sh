#!/bin/bash
# round 2, run 46 (1 GPU): attention core -- staging loads batched, row maximum over four threads per row, P.V unrolled by 8 -- parity, stack timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_transformer.py tests/test_gpu_dropin.py -m gpu -x -q -p no:cacheprovider > gpurun_out/r2_46_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_46_pytest.log | cut -c1-300
python - <<'PY' > gpurun_out/r2_46_stack.json 2> gpurun_out/r2_46_stack.err
import sys, json
sys.path.insert(0, "tools"); sys.path.insert(0, ".")
import bench_configs as bc
print(json.dumps(bc.attention_case()))
print(json.dumps(bc.transformer_case()))
PY
cat gpurun_out/r2_46_stack.json | cut -c1-400; tail -2 gpurun_out/r2_46_stack.err
python tools/prof_stack.py > gpurun_out/r2_46_prof_stack.json 2>&1; tail -1 gpurun_out/r2_46_prof_stack.json
