#!/bin/bash
# experimental stacked-dy weight gradient for C_out = 64 (ZNS_WGRAD_STACK=1): parity on the c_out=64 cases, then step time
export ZNS_WGRAD_STACK=1
timeout 60 python -m pytest tests/test_gpu_ops.py -x -q -k "(conv_wgrad_umma or conv_full) and 64-64" 2>&1 | tail -4 | tee gpurun_out/wstack_tests.txt
timeout 60 python bench.py --steps 20 --warmup 5 --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('stack', round(d['value'],1), round(d['ms_per_step'],3))" | tee gpurun_out/wstack_bench.txt
