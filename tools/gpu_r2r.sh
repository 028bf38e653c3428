#!/bin/bash
# VQT level 0: tensor-map staged loader vs register loader; parity tests; role counters
mkdir -p gpurun_out
L=$PWD/zeronotesamba_b200
timeout -k 10 600 python -m pytest tests -q -m gpu -p no:cacheprovider -k "vqt or xqt or config" > gpurun_out/r2r_tests.log 2>&1
echo "tests exit=$?"; tail -3 gpurun_out/r2r_tests.log | cut -c1-300
{
for rep in 1 2; do
echo "staged"; timeout 120 python tools/vqt_bench.py 20
echo "staged, no L2 prefetch"; ZNS_LIB_PATH=$L/libzns_sm100_nopf.so timeout 120 python tools/vqt_bench.py 20
echo "register path"; ZNS_VQT_TMA=0 timeout 120 python tools/vqt_bench.py 20
done
} 2>&1 | tee gpurun_out/r2r_vqt_ab.txt
ZNS_LIB_PATH=$L/libzns_sm100_timing.so timeout 120 python tools/vqt_bench.py 3 --timing 2>&1 | tail -9 > gpurun_out/r2r_role_counters.txt
cut -c1-200 gpurun_out/r2r_role_counters.txt | head -3
