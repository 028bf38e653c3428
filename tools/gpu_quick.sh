#!/bin/bash
# quick check of the tensor-core conv kernels: parity tests, then the step time (optionally per env setting)
set -o pipefail
timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -k "conv_fwd_umma or full_size or dgrad or two_branches or wgrad" 2>&1 | tail -5
rc=$?
echo "tests rc=$rc"
if [ $rc -eq 0 ]; then
  for i in 1 2 3; do
    timeout 200 python bench.py --steps 20 --warmup 5 --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1))"
  done
fi
