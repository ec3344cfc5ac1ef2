#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_conditioned_gpu.py -x -q -s > gpurun_out/r2_parity_cond.log 2>&1
timeout 600 python -m pytest tests/test_eval_gpu.py tests/test_fusion_gpu.py -q > gpurun_out/r2_eval_tests.log 2>&1
for p in 12 6; do UNETZOO_WGRAD_SM_PERCENT=$p python tools/step_time.py --tag wgrad_pct$p 2>/dev/null | tail -1 >> gpurun_out/r2_knobs2.log; done
UNETZOO_AUX_STREAMS=6 python tools/step_time.py --tag aux6 2>/dev/null | tail -1 >> gpurun_out/r2_knobs2.log
python tools/step_time.py --tag base 2>/dev/null | tail -1 >> gpurun_out/r2_knobs2.log
cat gpurun_out/r2_knobs2.log
grep -v "Warn\|warn" gpurun_out/r2_parity_cond.log | tail -40
tail -15 gpurun_out/r2_eval_tests.log
