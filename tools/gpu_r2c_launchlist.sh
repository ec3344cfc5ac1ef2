#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/r2c_launches_train_step.csv python tools/one_step.py 3 > gpurun_out/r2c_one_step.log 2>&1
tail -n 2 gpurun_out/r2c_one_step.log; wc -l gpurun_out/r2c_launches_train_step.csv
