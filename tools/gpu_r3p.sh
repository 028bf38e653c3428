#!/bin/bash
# CQT on the tcgen05 level kernels (n_fft 256): parity tests, big-batch determinism, agreement with the round-1 kernels
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests -q -m gpu -p no:cacheprovider -k "vqt or xqt or config" > gpurun_out/r3p_tests.log 2>&1
echo "tests exit=$?"; tail -3 gpurun_out/r3p_tests.log | cut -c1-300
timeout 300 python tools/cqt_check.py 2>&1 | tail -6 | tee gpurun_out/r3p_cqt.txt
timeout 120 python tools/vqt_bench.py 20
