#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_fusion_gpu.py tests/test_kernels_gpu.py -q -k "cluster or deferred or adam or kl_hierarchy or residual or slayer or head" 2>&1 | grep -v "Warning\|warn" | tail -15 > gpurun_out/r2c_memcheck.log; tail -6 gpurun_out/r2c_memcheck.log
