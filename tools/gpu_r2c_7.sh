#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_step_gpu.py tests/test_fusion_gpu.py tests/test_dp_gpu.py -q -x 2>&1 | tail -4
python tools/step_time.py --steps 60 --tag overlap_opt_geo 2>/dev/null | tail -1
UNETZOO_WGRAD_BIG_PERCENT=50 python tools/step_time.py --steps 60 --tag big50 2>/dev/null | tail -1
UNETZOO_WGRAD_BIG_PERCENT=100 python tools/step_time.py --steps 60 --tag big100 2>/dev/null | tail -1
UNETZOO_OVERLAP_OPT=0 python tools/step_time.py --steps 60 --tag no_overlap 2>/dev/null | tail -1
python tools/step_time.py --steps 60 --tag overlap_opt_geo_again 2>/dev/null | tail -1
