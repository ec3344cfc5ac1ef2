#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fusion_gpu.py tests/test_train_step_gpu.py tests/test_phiseg_gpu.py tests/test_dp_gpu.py tests/test_unet_probunet_gpu.py tests/test_phiseg3d_gpu.py tests/test_transparent_graph_gpu.py -q -x 2>&1 | tail -6
export UNETZOO_PRECISION=prof
python tools/step_time.py --multi-only --steps 40 --tag prof_base 2>/dev/null | tail -1
python tools/step_time.py --multi-only --steps 40 --debug-flags 256 --tag no_wgrad_kernels 2>/dev/null | tail -1
python tools/step_time.py --multi-only --steps 40 --debug-flags 4352 --tag no_wgrad_no_reduce 2>/dev/null | tail -1
python tools/step_time.py --multi-only --steps 40 --debug-flags 128 --tag no_conv 2>/dev/null | tail -1
python tools/step_time.py --multi-only --steps 40 --debug-flags 4480 --tag no_conv_no_wgrad 2>/dev/null | tail -1
