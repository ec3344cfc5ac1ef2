#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python tools/timeline.py --list conv_tc_kernel --out gpurun_out/r2c_tl_c.json 2>/dev/null | grep -E "grid|=="
UNETZOO_CONV_BN_FUSED=0 python tools/timeline.py --list conv_tc_kernel --out gpurun_out/r2c_tl_d.json 2>/dev/null | grep -E "grid|=="
