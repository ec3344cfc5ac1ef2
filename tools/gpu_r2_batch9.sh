#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
UNETZOO_PRECISION=prof timeout 120 python tools/phase_trace.py > gpurun_out/r2_phase_trace2.log 2>&1
cat gpurun_out/r2_phase_trace2.log | tail -80
timeout 1200 python -m pytest tests/test_fusion_gpu.py tests/test_kernels_gpu.py tests/test_phiseg_gpu.py tests/test_caller_contract_gpu.py -q -x > gpurun_out/r2_tests9.log 2>&1
tail -15 gpurun_out/r2_tests9.log
python tools/step_time.py --tag wgrad_epilogue 2>/dev/null | tail -1
python tools/step_time.py --model revphiseg --tag rev_fused 2>/dev/null | tail -1
UNETZOO_FUSED_REVERSIBLE=0 python tools/step_time.py --model revphiseg --tag rev_nested 2>/dev/null | tail -1
UNETZOO_WGRAD_SM_PERCENT=12 python tools/step_time.py --tag wgrad_pct12 2>/dev/null | tail -1
