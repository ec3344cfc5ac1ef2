#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_phiseg_gpu.py tests/test_kernels_gpu.py tests/test_train_step_gpu.py tests/test_caller_contract_gpu.py tests/test_eval_gpu.py tests/test_transparent_graph_gpu.py -q 2>&1 | grep -v "Warning\|warn" | grep -E "^[.sFE]+ *\[|FAILED|^E  |passed|failed" | head
python tools/step_time.py --steps 60 --multi-only --tag slayer_early 2>/dev/null | tail -1
python tools/step_time.py --steps 60 --multi-only --tag slayer_early 2>/dev/null | tail -1
