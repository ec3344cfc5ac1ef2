#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_eval_gpu.py tests/test_kernels_gpu.py tests/test_data_gpu.py -q 2>&1 | tail -4
python tools/eval_timeline.py 2>/dev/null > gpurun_out/r2_eval_timeline.txt; head -22 gpurun_out/r2_eval_timeline.txt
python tools/step_time.py --tag ew4 2>/dev/null | tail -1
UZ_EW_PER_THREAD=1 python tools/step_time.py --tag ew1 2>/dev/null | tail -1
UZ_EW_PER_THREAD=8 python tools/step_time.py --tag ew8 2>/dev/null | tail -1
