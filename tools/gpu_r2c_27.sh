#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
python tools/step_time.py --steps 60 --multi-only --tag pct25 2>/dev/null | tail -1
UNETZOO_WGRAD_SM_PERCENT=50 python tools/step_time.py --steps 60 --multi-only --tag pct50 2>/dev/null | tail -1
UNETZOO_WGRAD_SM_PERCENT=100 python tools/step_time.py --steps 60 --multi-only --tag pct100 2>/dev/null | tail -1
UNETZOO_WGRAD_BIG_PERCENT=100 python tools/step_time.py --steps 60 --multi-only --tag big100 2>/dev/null | tail -1
UNETZOO_WGRAD_SM_PERCENT=50 UNETZOO_WGRAD_BIG_PERCENT=100 python tools/step_time.py --steps 60 --multi-only --tag pct50big100 2>/dev/null | tail -1
python tools/step_time.py --steps 60 --multi-only --tag pct25 2>/dev/null | tail -1
