#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_step_gpu.py tests/test_kernels_gpu.py tests/test_data_gpu.py tests/test_fusion_gpu.py -q 2>&1 | grep -v "Warning\|warn" | grep -E "^[.sFE]+ *\[|FAILED|^E  " | head -20
for i in 1 2; do
python tools/step_time.py --steps 60 --multi-only --tag fromslabs 2>/dev/null | tail -1
UNETZOO_ADAM_FROM_SLABS=0 python tools/step_time.py --steps 60 --multi-only --tag reducepass 2>/dev/null | tail -1
done
python tools/step_time.py --steps 40 --multi-only --model revphiseg --tag rev 2>/dev/null | tail -1
