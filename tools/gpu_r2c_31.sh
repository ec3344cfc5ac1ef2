#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 300 python -m pytest tests/test_fusion_gpu.py -q -k "cooperative or cluster" 2>&1 | grep -v "Warning\|warn" | grep -E "^[.sFE]+ *\[|FAILED|^E  |passed|failed" | head
for i in 1 2; do
timeout 200 python tools/step_time.py --steps 60 --multi-only --tag coop 2>/dev/null | tail -1
UNETZOO_BN_BWD_COOP=0 timeout 200 python tools/step_time.py --steps 60 --multi-only --tag nocoop 2>/dev/null | tail -1
done
UNETZOO_BN_BWD_COOP_MIN_PIX=12288 timeout 200 python tools/step_time.py --steps 60 --multi-only --tag coop12288 2>/dev/null | tail -1
