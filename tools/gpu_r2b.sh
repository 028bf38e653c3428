#!/bin/bash
# round 2 session B: full -m gpu suite, then the per-role cycle counters of the VQT level kernels (timing build)
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r2b_tests.log 2>&1
echo "tests exit=$?" | tee gpurun_out/r2b_summary.txt
tail -25 gpurun_out/r2b_tests.log | cut -c1-220
grep -h "embedding error\|update sign\|cond grad\|cfg2 worst" gpurun_out/r2b_tests.log | head -40
ZNS_LIB_PATH=$PWD/zeronotesamba_b200/libzns_sm100_timing.so timeout 300 python tools/vqt_bench.py 3 --timing 2>&1 | tee gpurun_out/r2b_vqt_timing.txt | cut -c1-400
