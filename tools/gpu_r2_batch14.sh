#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_eval_gpu.py tests/test_kernels_gpu.py tests/test_phiseg_gpu.py tests/test_phiseg3d_gpu.py tests/test_kernels3d_gpu.py tests/test_unet_probunet_gpu.py -q 2>&1 | tail -4
python tools/eval_timeline.py 2>/dev/null > gpurun_out/r2_eval_timeline.txt; head -14 gpurun_out/r2_eval_timeline.txt
python tools/step_time.py --tag ew8_up2 2>/dev/null | tail -1
UZ_EW_PER_THREAD=16 python tools/step_time.py --tag ew16 2>/dev/null | tail -1
