#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python tools/timeline.py --multi-stream --list adam,pack_weight --out gpurun_out/r2c_tl_e.json 2>/dev/null | grep -E "grid|==|span"
