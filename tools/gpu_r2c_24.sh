#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fusion_gpu.py tests/test_kernels_gpu.py tests/test_kernels3d_gpu.py tests/test_phiseg_gpu.py tests/test_train_step_gpu.py tests/test_phiseg3d_gpu.py -q 2>&1 | grep -v "Warning\|warn" | grep -E "^[.sFE]+ *\[|FAILED|^E  " | head -20
python tools/step_time.py --steps 60 --multi-only --tag units8 2>/dev/null | tail -1
python tools/timeline.py --multi-stream --list wgrad_reduce,adam --out gpurun_out/r2c_tl_g.json 2>/dev/null | grep -E "grid|=="
