#!/bin/bash
# final checks of the round: full GPU suite, smoke(), bench N=1 (+ reference arm)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -v "Warning\|warn" > gpurun_out/r2c_full_tests5.log; grep -E "^[.sFE]+ *\[|FAILED|^E  " gpurun_out/r2c_full_tests5.log | head
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2c_bench_n1_c.json 2> gpurun_out/r2c_bench_n1_c.err; tail -c 200 gpurun_out/r2c_bench_n1_c.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2c_bench_ref.json 2> gpurun_out/r2c_bench_ref.err; tail -c 300 gpurun_out/r2c_bench_ref.json
