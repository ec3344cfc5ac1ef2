#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fusion_gpu.py tests/test_kernels_gpu.py tests/test_phiseg_gpu.py tests/test_train_step_gpu.py tests/test_parity_conditioned_gpu.py -q -x 2>&1 | grep -v "Warning\|warn" | tail -12
python tools/step_time.py --steps 60 --tag small 2>/dev/null | tail -1
UZ_CONV_SMALL=0 python tools/step_time.py --steps 60 --tag nosmall 2>/dev/null | tail -1
UZ_CONV_SMALL_PIX=64 python tools/step_time.py --steps 60 --multi-only --tag small64 2>/dev/null | tail -1
python tools/timeline.py --list conv_small --out gpurun_out/r2c_tl_f.json 2>/dev/null | grep -E "grid|=="
