#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dp_gpu.py tests/test_eval_gpu.py -q 2>&1 | grep -v "Warning\|warn" | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --skip-extra > gpurun_out/r2c_bench_n2.json 2> gpurun_out/r2c_bench_n2.err
tail -c 1500 gpurun_out/r2c_bench_n2.json; tail -3 gpurun_out/r2c_bench_n2.err
