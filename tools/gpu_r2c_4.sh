#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python tools/timeline.py --list bn_bwd --out gpurun_out/r2c_tl_a.json 2>/dev/null | grep -E "grid|==|span"
UNETZOO_BN_BWD_FUSED=0 python tools/timeline.py --list bn_bwd --out gpurun_out/r2c_tl_b.json 2>/dev/null | grep -E "grid|==|span"
