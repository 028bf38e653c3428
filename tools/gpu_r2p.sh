#!/bin/bash
# VQT: L2 prefetch distance A/B (1 / 2 / 3 tiles), ncu --set full capture of the deep levels (4..7)
mkdir -p gpurun_out
L=$PWD/zeronotesamba_b200
timeout -k 10 600 python -m pytest tests -q -m gpu -p no:cacheprovider -k "vqt or xqt" > gpurun_out/r2p_tests.log 2>&1
echo "tests exit=$?"; tail -3 gpurun_out/r2p_tests.log | cut -c1-200
{
for rep in 1 2; do
echo "dist 2"; python tools/vqt_bench.py 20
echo "dist 1"; ZNS_LIB_PATH=$L/libzns_sm100_d1.so python tools/vqt_bench.py 20
echo "dist 3"; ZNS_LIB_PATH=$L/libzns_sm100_d3.so python tools/vqt_bench.py 20
done
} 2>&1 | tee gpurun_out/r2p_vqt_ab.txt
ZNS_LIB_PATH=$L/libzns_sm100_timing.so python tools/vqt_bench.py 3 --timing 2>&1 | tail -9 > gpurun_out/r2p_role_counters.txt
cut -c1-200 gpurun_out/r2p_role_counters.txt | head -3
timeout -k 10 400 ncu --set full --clock-control none --import-source on -k regex:vqt_level_kernel -s 20 -c 4 -f -o gpurun_out/prof_vqt_r2p_deep python tools/vqt_bench.py 1 > gpurun_out/r2p_ncu.log 2>&1
tail -2 gpurun_out/r2p_ncu.log
