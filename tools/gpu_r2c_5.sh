#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_step_gpu.py tests/test_fusion_gpu.py tests/test_data_gpu.py tests/test_dp_gpu.py -q -x 2>&1 | tail -15
python tools/step_time.py --tag overlap_opt 2>/dev/null | tail -1
UNETZOO_OVERLAP_OPT=0 python tools/step_time.py --tag no_overlap 2>/dev/null | tail -1
