#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
for i in 1 2; do
python tools/step_time.py --steps 60 --multi-only --tag early 2>/dev/null | tail -1
UNETZOO_EARLY_LOGITS=0 python tools/step_time.py --steps 60 --multi-only --tag late 2>/dev/null | tail -1
done
