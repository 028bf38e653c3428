#!/bin/bash
# VQT: merged multi-frame accumulators (two stages + groups on levels 5 / 7), idle issuers out of the slot barrier; tests, A/B, counters
mkdir -p gpurun_out
L=$PWD/zeronotesamba_b200
timeout -k 10 600 python -m pytest tests -q -m gpu -p no:cacheprovider -k "vqt or xqt or config" > gpurun_out/r3c_tests.log 2>&1
echo "tests exit=$?"; tail -3 gpurun_out/r3c_tests.log | cut -c1-300
{
for rep in 1 2; do
for m in 6 2 0; do echo "groups=$m"; ZNS_VQT_GROUPS=$m timeout 120 python tools/vqt_bench.py 20; done
done
} 2>&1 | tee gpurun_out/r3c_vqt_ab.txt
ZNS_LIB_PATH=$L/libzns_sm100_timing.so timeout 120 python tools/vqt_bench.py 3 --timing 2>&1 | tail -5 | cut -c1-200
