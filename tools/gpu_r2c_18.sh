#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
python tools/step_time.py --steps 60 --multi-only --tag late32 2>/dev/null | tail -1
UNETZOO_WGRAD_LATE_MB=100000 python tools/step_time.py --steps 60 --multi-only --tag nolate 2>/dev/null | tail -1
python tools/step_time.py --steps 60 --multi-only --tag late32 2>/dev/null | tail -1
UNETZOO_WGRAD_LATE_MB=100000 python tools/step_time.py --steps 60 --multi-only --tag nolate 2>/dev/null | tail -1
UNETZOO_WGRAD_LATE_PIXELS=12288 python tools/step_time.py --steps 60 --multi-only --tag late32_32sq 2>/dev/null | tail -1
