#!/bin/bash
# round 2, run 53 (1 GPU): 128-column tiles for products that fill at most half the machine -- whole GPU suite, stack timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r2_53_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_53_pytest.log | cut -c1-300
python - <<'PY' > gpurun_out/r2_53_stack.json 2> gpurun_out/r2_53_stack.err
import sys, json
sys.path.insert(0, "tools"); sys.path.insert(0, ".")
import bench_configs as bc
print(json.dumps(bc.transformer_case()))
PY
cat gpurun_out/r2_53_stack.json | cut -c1-400; tail -2 gpurun_out/r2_53_stack.err
QG_NO_SMALL_HALF=1 python - <<'PY' > gpurun_out/r2_53_stack_off.json 2>> gpurun_out/r2_53_stack.err
import sys, json
sys.path.insert(0, "tools"); sys.path.insert(0, ".")
import bench_configs as bc
print(json.dumps(bc.transformer_case()))
PY
cat gpurun_out/r2_53_stack_off.json | cut -c1-400
