#!/bin/bash
# ncu --set full captures of one warm launch per kernel family; only CSV summaries travel back
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 500 $NCU -k regex:'conv_tc|wgrad_tc|wgrad_reduce|bn_' -o /tmp/p_conv python tools/ncu_probe.py --what conv_big,conv_small,conv_tiny,conv_c32,wgrad_mid,wgrad_small,wgrad_tiny,bn_c128,bn_c32 > gpurun_out/r2b_probe_conv.log 2>&1
ncu -i /tmp/p_conv.ncu-rep --page raw --csv > gpurun_out/r2b_probe_conv_raw.csv 2>/dev/null
timeout 400 $NCU -k regex:'eval_|pair_distance|ged_finish|ncc_|dice_|pack_masks|head_|slayer_|residual_ce|kl_|avgpool|up2_|column_reduce' -o /tmp/p_mem python tools/ncu_probe.py --what eval_tail,heads,memops > gpurun_out/r2b_probe_mem.log 2>&1
ncu -i /tmp/p_mem.ncu-rep --page raw --csv > gpurun_out/r2b_probe_mem_raw.csv 2>/dev/null
# launch list of one graph-replayed training step (cold-cache serialised durations: compare SHARES)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2b_launches_train_step.csv python tools/one_step.py > gpurun_out/r2b_one_step.log 2>&1
ls -la /tmp/*.ncu-rep gpurun_out/ | tail -20
