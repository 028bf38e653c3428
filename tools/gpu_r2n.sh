#!/bin/bash
# VQT A/B: loader chunks in flight (4 / 6 / 8), programmatic dependent launch on / off; parity tests; trainer graphs with PDL
mkdir -p gpurun_out
L=$PWD/zeronotesamba_b200
timeout -k 10 600 python -m pytest tests -q -m gpu -p no:cacheprovider -k "vqt or xqt or config or trainer or prefetch" > gpurun_out/r2n_tests.log 2>&1
echo "tests exit=$?"; tail -3 gpurun_out/r2n_tests.log | cut -c1-200
{
for rep in 1 2; do
echo "U=6 PDL on";  python tools/vqt_bench.py 20
echo "U=6 PDL off"; ZNS_VQT_PDL=0 python tools/vqt_bench.py 20
echo "U=4 PDL on";  ZNS_LIB_PATH=$L/libzns_sm100_u4.so python tools/vqt_bench.py 20
echo "U=8 PDL on";  ZNS_LIB_PATH=$L/libzns_sm100_u8.so python tools/vqt_bench.py 20
done
} 2>&1 | tee gpurun_out/r2n_vqt_ab.txt
ZNS_LIB_PATH=$L/libzns_sm100_timing.so python tools/vqt_bench.py 3 --timing 2>&1 | tail -9 > gpurun_out/r2n_role_counters.txt
cut -c1-200 gpurun_out/r2n_role_counters.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1))"
