#!/bin/bash
# evidence run: cfg2 VQT launch list with DRAM traffic, full tests, default bench line, conv launch list + traffic
mkdir -p gpurun_out
timeout -k 10 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 18 -c 9 --csv --log-file gpurun_out/r02_vqt_launches.csv python tools/vqt_bench.py 1 > /dev/null 2>&1
python tools/vqt_bench.py 5 | tee gpurun_out/r02_vqt_bench.txt
timeout -k 10 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r2k_tests.log 2>&1
echo "tests exit=$?"; tail -3 gpurun_out/r2k_tests.log | cut -c1-200
timeout 900 python bench.py > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo "bench exit=$?"
timeout -k 10 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_step_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --sustained-s 0 > /dev/null 2>&1
python tools/ncu_launch_shares.py gpurun_out/r02_step_launches.csv "r02" | head -12
