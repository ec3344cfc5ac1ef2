#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fusion_gpu.py tests/test_kernels_gpu.py -q -x 2>&1 | tail -15
python tools/step_time.py --tag bn_cluster 2>/dev/null | tail -1
UNETZOO_BN_BWD_FUSED=0 python tools/step_time.py --tag bn_unfused 2>/dev/null | tail -1
