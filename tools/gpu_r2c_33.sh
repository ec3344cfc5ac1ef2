#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 300 python -m pytest tests/test_eval_gpu.py tests/test_phiseg_gpu.py -q 2>&1 | grep -v "Warning\|warn" | grep -E "^[.sFE]+ *\[|FAILED|^E  |passed|failed" | head
python bench.py --skip-cpu --skip-torch --skip-extra --steps 10 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['eval_ged100']; print('train', d['value'], d['ms_per_step'], 'eval', e['value'], e['ms_per_call'], 'static', e['value_static_weights'], e['ms_per_call_static_weights'])"
