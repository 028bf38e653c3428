#!/bin/bash
# A/B on one box: stacked-dy weight gradient for cv2 (ZNS_WGRAD_STACK=1) vs the default N=64 kernel; three alternations
for i in 1 2 3; do
  for v in 0 1; do
    echo "ZNS_WGRAD_STACK=$v: $(ZNS_WGRAD_STACK=$v timeout 300 python bench.py --no-extras --sustained-s 0 --steps 100 --warmup 10 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],1), round(d['ms_per_step'],4))")"
  done
done
ZNS_WGRAD_STACK=1 timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -p no:cacheprovider -k "conv_wgrad_umma or conv_full" 2>&1 | tail -2
