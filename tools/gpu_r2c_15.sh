#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_train_step_gpu.py tests/test_kernels3d_gpu.py -q -x 2>&1 | tail -3
python tools/timeline.py --multi-stream --list adam,pack_weight --out gpurun_out/r2c_tl_e.json 2>/dev/null | grep -E "grid|==|span"
python tools/step_time.py --steps 60 --multi-only --tag adampack 2>/dev/null | tail -1
UNETZOO_ADAM_PACK=0 python tools/step_time.py --steps 60 --multi-only --tag no_adampack 2>/dev/null | tail -1
python tools/step_time.py --steps 60 --multi-only --tag adampack 2>/dev/null | tail -1
UNETZOO_ADAM_PACK=0 python tools/step_time.py --steps 60 --multi-only --tag no_adampack 2>/dev/null | tail -1
