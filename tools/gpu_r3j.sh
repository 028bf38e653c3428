#!/bin/bash
# VQT level 0: L2 prefetch of the next tile issued before the fill by all lanes (early) vs behind the fill by one lane
mkdir -p gpurun_out
L=$PWD/zeronotesamba_b200
ZNS_LIB_PATH=$L/libzns_sm100_pfearly.so timeout 300 python tools/vqt_diff.py $L/libzns_sm100.so 2>&1 | grep "level\|run-to" | tail -10 | cut -c1-80
{
for rep in 1 2 3; do
echo "early prefetch"; ZNS_LIB_PATH=$L/libzns_sm100_pfearly.so timeout 120 python tools/vqt_bench.py 20
echo "default"; timeout 120 python tools/vqt_bench.py 20
done
} 2>&1 | tee gpurun_out/r3j_vqt_ab.txt
