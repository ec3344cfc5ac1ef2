#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 --skip-extra > gpurun_out/r2c_bench_n8.json 2> gpurun_out/r2c_bench_n8.err
tail -c 400 gpurun_out/r2c_bench_n8.json; tail -2 gpurun_out/r2c_bench_n8.err | cut -c1-300
