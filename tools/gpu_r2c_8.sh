#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fusion_gpu.py tests/test_train_step_gpu.py tests/test_phiseg_gpu.py tests/test_dp_gpu.py tests/test_unet_probunet_gpu.py tests/test_phiseg3d_gpu.py tests/test_transparent_graph_gpu.py -q -x 2>&1 | tail -6
python tools/step_time.py --steps 60 --tag defer 2>/dev/null | tail -1
UNETZOO_DEFER_WGRAD_REDUCE=0 python tools/step_time.py --steps 60 --tag nodefer 2>/dev/null | tail -1
UNETZOO_WGRAD_FLUSH_MB=16 python tools/step_time.py --steps 60 --tag defer16 2>/dev/null | tail -1
UNETZOO_WGRAD_FLUSH_MB=4096 python tools/step_time.py --steps 60 --tag defer_end 2>/dev/null | tail -1
