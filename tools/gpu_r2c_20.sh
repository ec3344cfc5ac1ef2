#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -v "Warning\|warn" > gpurun_out/r2c_full_tests4.log; grep -E "^[.sFE]+ *\[|FAILED|^E  " gpurun_out/r2c_full_tests4.log | head -20
python tools/step_time.py --steps 60 --multi-only --tag small 2>/dev/null | tail -1
UZ_CONV_SMALL=0 python tools/step_time.py --steps 60 --multi-only --tag nosmall 2>/dev/null | tail -1
python tools/step_time.py --steps 60 --multi-only --tag small 2>/dev/null | tail -1
UZ_CONV_SMALL=0 python tools/step_time.py --steps 60 --multi-only --tag nosmall 2>/dev/null | tail -1
