#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
UNETZOO_PRECISION=prof timeout 120 python tools/phase_trace.py > gpurun_out/r2_phase_trace.log 2>&1
cat gpurun_out/r2_phase_trace.log | tail -50
timeout 900 python -m pytest tests/test_caller_contract_gpu.py tests/test_baseline_configs_gpu.py tests/test_parity_conditioned_gpu.py -q -s > gpurun_out/r2_contract_tests.log 2>&1
grep -n "validate()\|sample()\|U-Net B=12\|ProbUNet B=12\|PHISeg3D\|passed\|failed\|Error" gpurun_out/r2_contract_tests.log | head -30
