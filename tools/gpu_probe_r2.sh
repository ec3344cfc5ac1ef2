#!/bin/bash
# ncu captures of the small-shape tensor-core kernels and the BatchNorm passes; only CSV summaries travel back
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:'wgrad_tc_kernel|conv_tc_kernel' -o /tmp/p_tiny python tools/ncu_probe.py --what wgrad_tiny,conv_tiny > gpurun_out/r2_probe_tiny.log 2>&1
ncu -i /tmp/p_tiny.ncu-rep --page raw --csv > gpurun_out/r2_probe_tiny_raw.csv 2>/dev/null
ncu -i /tmp/p_tiny.ncu-rep --page source --csv > gpurun_out/r2_probe_tiny_source.csv 2>/dev/null
timeout 400 $NCU -k regex:'wgrad_tc2|conv_tc2|wgrad_reduce|bn_' -o /tmp/p_mid python tools/ncu_probe.py --what wgrad_small,wgrad_mid,conv_small,conv_c32,conv_big,bn_c32,bn_c128 > gpurun_out/r2_probe_mid.log 2>&1
ncu -i /tmp/p_mid.ncu-rep --page raw --csv > gpurun_out/r2_probe_mid_raw.csv 2>/dev/null
ls -la /tmp/*.ncu-rep gpurun_out/
