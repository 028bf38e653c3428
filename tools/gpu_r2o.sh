#!/bin/bash
# VQT: unified level-0 loader path; parity tests; cfg2 time; ncu --set full capture of level 0 and level 1
mkdir -p gpurun_out
L=$PWD/zeronotesamba_b200
timeout -k 10 600 python -m pytest tests -q -m gpu -p no:cacheprovider -k "vqt or xqt or config" > gpurun_out/r2o_tests.log 2>&1
echo "tests exit=$?"; tail -3 gpurun_out/r2o_tests.log | cut -c1-200
for rep in 1 2; do python tools/vqt_bench.py 20; done 2>&1 | tee gpurun_out/r2o_vqt.txt
ZNS_LIB_PATH=$L/libzns_sm100_timing.so python tools/vqt_bench.py 3 --timing 2>&1 | tail -9 > gpurun_out/r2o_role_counters.txt
cut -c1-200 gpurun_out/r2o_role_counters.txt
timeout -k 10 400 ncu --set full --clock-control none --import-source on -k regex:vqt_level_kernel -s 16 -c 2 -f -o gpurun_out/prof_vqt_r2o python tools/vqt_bench.py 1 > gpurun_out/r2o_ncu.log 2>&1
tail -2 gpurun_out/r2o_ncu.log
