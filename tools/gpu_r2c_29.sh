#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
python tools/step_time.py --steps 60 --multi-only --tag base 2>/dev/null | tail -1
UNETZOO_FUSE_BN_BWD=1 python tools/step_time.py --steps 60 --multi-only --tag fuse_bn_bwd 2>/dev/null | tail -1
UNETZOO_AUX_STREAMS=4 python tools/step_time.py --steps 60 --multi-only --tag aux4 2>/dev/null | tail -1
UNETZOO_AUX_STREAMS=2 python tools/step_time.py --steps 60 --multi-only --tag aux2 2>/dev/null | tail -1
UZ_EW_PER_THREAD=8 python tools/step_time.py --steps 60 --multi-only --tag ew8 2>/dev/null | tail -1
python tools/step_time.py --steps 60 --multi-only --tag base 2>/dev/null | tail -1
