#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
UZ_PDL=3 python tools/step_time.py --steps 60 --tag pdl3 2>/dev/null | tail -1
UZ_PDL=1 python tools/step_time.py --steps 60 --tag pdl1 2>/dev/null | tail -1
python tools/step_time.py --steps 60 --multi-only --tag pdl0 2>/dev/null | tail -1
UZ_PDL=3 python tools/step_time.py --steps 60 --multi-only --tag pdl3 2>/dev/null | tail -1
UZ_PDL=1 python tools/step_time.py --steps 60 --multi-only --tag pdl1 2>/dev/null | tail -1
