#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
UNETZOO_PRECISION=prof timeout 120 python tools/phase_trace.py 2>&1 | grep -A14 "^wgrad 192->192 @4x4"
timeout 1500 python -m pytest tests/test_fusion_gpu.py tests/test_kernels_gpu.py tests/test_phiseg_gpu.py tests/test_caller_contract_gpu.py tests/test_unet_probunet_gpu.py -q > gpurun_out/r2_tests10.log 2>&1
tail -15 gpurun_out/r2_tests10.log
python tools/step_time.py --tag wgrad_chunks 2>/dev/null | tail -1
python tools/layer_times.py > gpurun_out/r2_layer_times2.log 2>&1; grep "^{" gpurun_out/r2_layer_times2.log
