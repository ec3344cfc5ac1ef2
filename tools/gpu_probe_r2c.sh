#!/bin/bash
# ncu --set full captures of one warm launch per kernel family (final round-2 code); only CSV summaries travel back
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 500 $NCU -k regex:'conv_tc|wgrad_tc|wgrad_reduce|bn_|adam_' -o /tmp/p_conv python tools/ncu_probe.py --what conv_big,conv_small,conv_tiny,conv_c32,fused_8x8,fused_2x2,wgrad_mid,wgrad_small,wgrad_tiny,bn_c128,bn_c32,bn_cluster_16,bn_cluster_4,optimizer > gpurun_out/r2c_probe_conv.log 2>&1
ncu -i /tmp/p_conv.ncu-rep --page raw --csv > gpurun_out/r2c_probe_conv_raw.csv 2>/dev/null
timeout 400 $NCU -k regex:'eval_|pair_distance|ged_finish|ncc_|dice_|pack_masks|head_|slayer_|residual_ce|kl_|avgpool|up2_|column_reduce' -o /tmp/p_mem python tools/ncu_probe.py --what eval_tail,heads,memops,kl_hier > gpurun_out/r2c_probe_mem.log 2>&1
ncu -i /tmp/p_mem.ncu-rep --page raw --csv > gpurun_out/r2c_probe_mem_raw.csv 2>/dev/null
# launch list of eager training steps (cold-cache serialised durations: compare SHARES)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/r2c_launches_train_step.csv python tools/one_step.py 3 > gpurun_out/r2c_one_step.log 2>&1
tail -2 gpurun_out/r2c_probe_conv.log gpurun_out/r2c_probe_mem.log gpurun_out/r2c_one_step.log
ls -la gpurun_out/ | grep r2c_probe
