#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -v "Warning\|warn" > gpurun_out/r2c_full_tests3.log; grep -c "passed\|failed" gpurun_out/r2c_full_tests3.log; tail -3 gpurun_out/r2c_full_tests3.log | cut -c1-150
timeout 900 python bench.py > gpurun_out/r2c_bench_n1_b.json 2> gpurun_out/r2c_bench_n1_b.err; tail -c 300 gpurun_out/r2c_bench_n1_b.json
