#!/bin/bash
# bias gradient with four loads in flight; cv1 forward with 2 / 3 / 4 frames per thread (kernel times under ncu, step time)
mkdir -p gpurun_out
L=$PWD/zeronotesamba_b200
timeout -k 10 600 python -m pytest tests -q -m gpu -p no:cacheprovider -k "bias_grad or conv1 or two_branches or golden or cfg3" > gpurun_out/r2z_tests.log 2>&1
echo "tests exit=$?"; tail -3 gpurun_out/r2z_tests.log | cut -c1-300
for v in default npos3 npos4; do
  lib=$L/libzns_sm100.so; [ $v != default ] && lib=$L/libzns_sm100_$v.so
  ZNS_LIB_PATH=$lib timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"conv1_fwd_kernel|bias_grad_kernel" -s 16 -c 8 --csv --log-file gpurun_out/r2z_$v.csv python bench.py --steps 1 --warmup 3 --no-extras --sustained-s 0 > /dev/null 2>&1
  echo "$v: $(grep 'conv1_fwd\|bias_grad' gpurun_out/r2z_$v.csv | awk -F'","' '{n=split($5,a,"("); printf "%s=%s ", a[1], $NF}' | tr -d '"')"
done
for v in default npos3 npos4 default; do
  lib=$L/libzns_sm100.so; [ $v != default ] && lib=$L/libzns_sm100_$v.so
  ZNS_LIB_PATH=$lib timeout 200 python bench.py --steps 30 --warmup 5 --no-extras --sustained-s 0 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v bench', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1))"
done
