#!/bin/bash
# VQT: one mbarrier arrive per warp vs per lane
mkdir -p gpurun_out
L=$PWD/zeronotesamba_b200
timeout -k 10 600 python -m pytest tests -q -m gpu -p no:cacheprovider -k "vqt or xqt or config" > gpurun_out/r2x_tests.log 2>&1
echo "tests exit=$?"; tail -3 gpurun_out/r2x_tests.log | cut -c1-300
{
for rep in 1 2; do
echo "arrive per warp"; timeout 120 python tools/vqt_bench.py 20
echo "arrive per lane"; ZNS_LIB_PATH=$L/libzns_sm100_lanearrive.so timeout 120 python tools/vqt_bench.py 20
done
} 2>&1 | tee gpurun_out/r2x_vqt_ab.txt
ZNS_LIB_PATH=$L/libzns_sm100_timing.so timeout 120 python tools/vqt_bench.py 3 --timing 2>&1 | tail -9 > gpurun_out/r2x_role_counters.txt
cut -c1-200 gpurun_out/r2x_role_counters.txt
