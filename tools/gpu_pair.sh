#!/bin/bash
# CTA-pair (cta_group::2) conv kernels: parity tests, then step time per mode
#   mode 0 = single-CTA kernels, 1 = N=128 + stacked fwd/dgrad, 2 = + N=256, 3 = + weight gradient
set -o pipefail
TOP=${1:-3}
ZNS_CONV_PAIR=$TOP timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -k "conv_fwd_umma or full_size or dgrad or two_branches or wgrad" 2>&1 | tail -15
rc=$?
echo "tests rc=$rc"
if [ $rc -eq 0 ]; then
  for i in 1 2; do
    for m in $((TOP-1)) $TOP; do
      ZNS_CONV_PAIR=$m timeout 200 python bench.py --steps 20 --warmup 5 --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('mode $m', round(d['value'],1), round(d['ms_per_step'],3))"
    done
  done
fi
