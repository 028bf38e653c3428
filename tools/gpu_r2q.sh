#!/bin/bash
# VQT: 16 epilogue warps + single bulk-copy loader warp for levels >= 1; parity tests; cfg2 time; role counters
mkdir -p gpurun_out
L=$PWD/zeronotesamba_b200
timeout -k 10 600 python -m pytest tests -q -m gpu -p no:cacheprovider -k "vqt or xqt or config" > gpurun_out/r2q_tests.log 2>&1
echo "tests exit=$?"; tail -3 gpurun_out/r2q_tests.log | cut -c1-200
for rep in 1 2; do python tools/vqt_bench.py 20; done 2>&1 | tee gpurun_out/r2q_vqt.txt
ZNS_LIB_PATH=$L/libzns_sm100_timing.so python tools/vqt_bench.py 3 --timing 2>&1 | tail -9 > gpurun_out/r2q_role_counters.txt
cut -c1-200 gpurun_out/r2q_role_counters.txt
