#!/bin/bash
mkdir -p gpurun_out
L=$PWD/zeronotesamba_b200
ZNS_LIB_PATH=$L/libzns_sm100_decfirst.so timeout 300 python tools/vqt_diff.py $L/libzns_sm100_prev.so 2>&1 | grep "level\|run-to" | tail -10
{
for rep in 1 2 3; do
echo "fb ring first"; timeout 120 python tools/vqt_bench.py 20
echo "dec ring first"; ZNS_LIB_PATH=$L/libzns_sm100_decfirst.so timeout 120 python tools/vqt_bench.py 20
echo "prev"; ZNS_LIB_PATH=$L/libzns_sm100_prev.so timeout 120 python tools/vqt_bench.py 20
done
} 2>&1 | tee gpurun_out/r3e_vqt_ab.txt
