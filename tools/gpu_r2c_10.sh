#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_fusion_gpu.py -q -x -k "cluster or deferred" 2>&1 | tail -12
timeout 900 python -m pytest tests/test_train_step_gpu.py tests/test_phiseg_gpu.py tests/test_unet_probunet_gpu.py tests/test_parity_conditioned_gpu.py -q -x 2>&1 | tail -6
python tools/step_time.py --steps 60 --tag convbn 2>/dev/null | tail -1
UNETZOO_CONV_BN_FUSED=0 python tools/step_time.py --steps 60 --tag no_convbn 2>/dev/null | tail -1
