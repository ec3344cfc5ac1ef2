#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python tools/timeline.py --multi-stream --raw gpurun_out/r2c_raw_ms.json --out gpurun_out/r2c_timeline_ms.json 2>/dev/null | tail -3
python tools/timeline.py --raw gpurun_out/r2c_raw_ss.json --out gpurun_out/r2c_timeline_ss.json 2>/dev/null | tail -3
