#!/bin/bash
# VQT levels >= 1: issuer tables (packed MMAs, segments) read from shared memory instead of the parameter bank; diff, tests, A/B
mkdir -p gpurun_out
L=$PWD/zeronotesamba_b200
timeout 300 python tools/vqt_diff.py $L/libzns_sm100_prev.so 2>&1 | grep "level\|run-to" | tail -10 | tee gpurun_out/r3i_diff.txt
timeout -k 10 600 python -m pytest tests -q -m gpu -p no:cacheprovider -k "vqt or xqt or config" > gpurun_out/r3i_tests.log 2>&1
echo "tests exit=$?"; tail -3 gpurun_out/r3i_tests.log | cut -c1-300
{
for rep in 1 2 3; do
echo "issuer tables in shared memory"; timeout 120 python tools/vqt_bench.py 20
echo "prev"; ZNS_LIB_PATH=$L/libzns_sm100_prev.so timeout 120 python tools/vqt_bench.py 20
done
} 2>&1 | tee gpurun_out/r3i_vqt_ab.txt
ZNS_LIB_PATH=$L/libzns_sm100_timing.so timeout 120 python tools/vqt_bench.py 3 --timing 2>&1 | tail -6 | cut -c1-200
