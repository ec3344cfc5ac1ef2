#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_conditioned_gpu.py -q -s > gpurun_out/r2_parity_cond2.log 2>&1
grep -n "storage, conditioned\|ELBO\|gradient rel\|passed\|failed\|Error\|error" gpurun_out/r2_parity_cond2.log | head -30
timeout 600 python -m pytest tests/test_eval_gpu.py tests/test_dp_gpu.py -q -x > gpurun_out/r2_n2_tests2.log 2>&1
tail -8 gpurun_out/r2_n2_tests2.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1_a.json 2> gpurun_out/r2_bench_n1_a.err
tail -c 3000 gpurun_out/r2_bench_n1_a.json; tail -5 gpurun_out/r2_bench_n1_a.err
