#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python tools/timeline.py --multi-stream --raw gpurun_out/r2c_raw_ms2.json --out gpurun_out/r2c_timeline_ms2.json 2>/dev/null | tail -2
