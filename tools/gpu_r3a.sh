#!/bin/bash
# VQT multi-frame levels: x2.g1 accumulated onto the x1.g2 columns (two accumulator stages + epilogue groups on levels 5 and 7)
mkdir -p gpurun_out
L=$PWD/zeronotesamba_b200
timeout -k 10 600 python -m pytest tests -q -m gpu -p no:cacheprovider -k "vqt or xqt or config" > gpurun_out/r3a_tests.log 2>&1
echo "tests exit=$?"; tail -3 gpurun_out/r3a_tests.log | cut -c1-300
{
for rep in 1 2; do
for m in 2 0; do echo "groups=$m"; ZNS_VQT_GROUPS=$m timeout 120 python tools/vqt_bench.py 20; done
done
} 2>&1 | tee gpurun_out/r3a_vqt_ab.txt
ZNS_LIB_PATH=$L/libzns_sm100_timing.so timeout 120 python tools/vqt_bench.py 3 --timing 2>&1 | tail -5 | cut -c1-200
