#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -v "Warning\|warn" > gpurun_out/r2c_full_tests2.log; tail -4 gpurun_out/r2c_full_tests2.log
python tools/step_time.py --steps 60 --tag adampack 2>/dev/null | tail -1
UNETZOO_ADAM_PACK=0 python tools/step_time.py --steps 60 --tag no_adampack 2>/dev/null | tail -1
