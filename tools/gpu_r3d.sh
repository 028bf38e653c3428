#!/bin/bash
# VQT: four accumulator stages + four epilogue groups on the two-frame levels; diff against the previous build, tests, A/B
mkdir -p gpurun_out
L=$PWD/zeronotesamba_b200
for m in 6 0; do echo "== groups=$m"; ZNS_VQT_GROUPS=$m timeout 300 python tools/vqt_diff.py $L/libzns_sm100_prev.so 2>&1 | grep -v "clips:\|distinct\|^ \[\|^  *[0-9]" | tail -12; done | tee gpurun_out/r3d_diff.txt
timeout -k 10 600 python -m pytest tests -q -m gpu -p no:cacheprovider -k "vqt or xqt or config" > gpurun_out/r3d_tests.log 2>&1
echo "tests exit=$?"; tail -3 gpurun_out/r3d_tests.log | cut -c1-300
{
for rep in 1 2; do
echo "new groups=6"; timeout 120 python tools/vqt_bench.py 20
echo "new groups=0"; ZNS_VQT_GROUPS=0 timeout 120 python tools/vqt_bench.py 20
echo "prev"; ZNS_LIB_PATH=$L/libzns_sm100_prev.so timeout 120 python tools/vqt_bench.py 20
done
} 2>&1 | tee gpurun_out/r3d_vqt_ab.txt
ZNS_LIB_PATH=$L/libzns_sm100_timing.so timeout 120 python tools/vqt_bench.py 3 --timing 2>&1 | tail -5 | cut -c1-200
