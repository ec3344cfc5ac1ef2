#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r2c_tests.log; tail -3 gpurun_out/r2c_tests.log
timeout 900 python bench.py > gpurun_out/r2c_bench_n1.json 2> gpurun_out/r2c_bench_n1.err; tail -c 600 gpurun_out/r2c_bench_n1.json
