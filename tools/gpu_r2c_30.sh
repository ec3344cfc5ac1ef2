#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
for mb in 2 3 4 2 3; do
UZ_ADAM_PACK_MINB=$mb python tools/step_time.py --steps 60 --multi-only --tag minb$mb 2>/dev/null | tail -1
done
