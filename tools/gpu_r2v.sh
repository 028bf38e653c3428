#!/bin/bash
# cv1 forward with two frames per thread: parity tests, kernel time under ncu, step time
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests -q -m gpu -p no:cacheprovider -k "conv1 or two_branches or golden or trainer or cfg3" > gpurun_out/r2v_tests.log 2>&1
echo "tests exit=$?"; tail -3 gpurun_out/r2v_tests.log | cut -c1-300
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"conv1_fwd_kernel|conv1_wgrad_kernel" -s 4 -c 4 --csv --log-file gpurun_out/r2v_cv1.csv python bench.py --steps 1 --warmup 3 --no-extras --sustained-s 0 > /dev/null 2>&1
grep -o '"conv1_[a-z_]*kernel.*' gpurun_out/r2v_cv1.csv | sed 's/(C1[^"]*"/"/' | cut -c1-120
for i in 1 2 3; do
  timeout 200 python bench.py --steps 20 --warmup 5 --no-extras --sustained-s 0 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1))"
done
