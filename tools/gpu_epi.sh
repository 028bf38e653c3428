#!/bin/bash
# A/B: epilogue warps per TMEM lane quadrant (2 = default build, 3 and 4 = side builds)
for i in 1 2; do
for e in 2 3 4; do
  lib=zeronotesamba_b200/libzns_e$e.so; [ $e -eq 2 ] && lib=zeronotesamba_b200/libzns_sm100.so
  ZNS_LIB_PATH=$PWD/$lib timeout 200 python bench.py --steps 20 --warmup 5 --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('EPI=$e', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1))"
done; done
ZNS_LIB_PATH=$PWD/zeronotesamba_b200/libzns_e4.so timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -k "conv_fwd_umma or full_size or dgrad or two_branches or wgrad" 2>&1 | tail -3
