#!/bin/bash
# VQT: rewritten edge-frame kernel; parity tests; cfg2 time; per-launch times
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests -q -m gpu -p no:cacheprovider -k "vqt or xqt or config" > gpurun_out/r2t_tests.log 2>&1
echo "tests exit=$?"; tail -3 gpurun_out/r2t_tests.log | cut -c1-300
for rep in 1 2; do timeout 120 python tools/vqt_bench.py 20; done 2>&1 | tee gpurun_out/r2t_vqt.txt
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 18 -c 9 --csv --log-file gpurun_out/r2t_vqt_launches.csv python tools/vqt_bench.py 1 > /dev/null 2>&1
grep -o '"vqt_[a-z_]*kernel[^"]*".*' gpurun_out/r2t_vqt_launches.csv | cut -c1-60,160-260 | head -12
