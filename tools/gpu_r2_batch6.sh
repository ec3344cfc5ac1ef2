#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dp_gpu.py tests/test_eval_gpu.py -q -x > gpurun_out/r2_n2_tests.log 2>&1
tail -12 gpurun_out/r2_n2_tests.log
timeout 900 python -m pytest tests/test_parity_conditioned_gpu.py -q -s > gpurun_out/r2_parity_cond.log 2>&1
grep -n "conditioned\|ELBO\|levels\|passed\|failed" gpurun_out/r2_parity_cond.log | head -30
timeout 300 python -m pytest tests/test_fusion_gpu.py -q -s 2>&1 | tail -5
