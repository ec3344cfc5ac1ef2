#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 600 python -m pytest tests/test_train_step_gpu.py tests/test_kernels_gpu.py -q 2>&1 | grep -v "Warning\|warn" | grep -E "^[.sFE]+ *\[|FAILED|^E  " | head -20
for i in 1 2; do
python tools/step_time.py --steps 60 --multi-only --tag fromslabs 2>/dev/null | tail -1
UNETZOO_ADAM_FROM_SLABS=0 python tools/step_time.py --steps 60 --multi-only --tag reducepass 2>/dev/null | tail -1
done
python tools/timeline.py --multi-stream --list adam --out gpurun_out/r2c_tl_h.json 2>/dev/null | grep -E "grid|=="
