#!/bin/bash
# VQT check: parity tests of the front-end, cfg2 time, role counters of the timing build
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests -q -m gpu -p no:cacheprovider -k "vqt or xqt or config" > gpurun_out/r2m_tests.log 2>&1
echo "tests exit=$?"; tail -3 gpurun_out/r2m_tests.log | cut -c1-200
for i in 1 2 3; do python tools/vqt_bench.py 10; done | tee gpurun_out/r2m_vqt_bench.txt
ZNS_LIB_PATH=$PWD/zeronotesamba_b200/libzns_sm100_timing.so python tools/vqt_bench.py 3 --timing 2>&1 | tail -9 > gpurun_out/r2m_role_counters.txt
cat gpurun_out/r2m_role_counters.txt | cut -c1-250
