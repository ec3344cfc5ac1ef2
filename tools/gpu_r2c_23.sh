#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fusion_gpu.py tests/test_kernels_gpu.py tests/test_kernels3d_gpu.py tests/test_phiseg_gpu.py tests/test_train_step_gpu.py tests/test_dp_gpu.py tests/test_phiseg3d_gpu.py tests/test_parity_conditioned_gpu.py -q 2>&1 | grep -v "Warning\|warn" | grep -E "^[.sFE]+ *\[|FAILED|^E  " | head -20
for i in 1 2; do
python tools/step_time.py --steps 60 --multi-only --tag acc 2>/dev/null | tail -1
UNETZOO_WGRAD_ACCUMULATE=0 python tools/step_time.py --steps 60 --multi-only --tag slabs 2>/dev/null | tail -1
done
