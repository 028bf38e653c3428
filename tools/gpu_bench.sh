#!/bin/bash
# Model-level parity tests, the bench line, and the ncu launch list + one full capture of the top kernel.
mkdir -p gpurun_out
: > gpurun_out/summary.txt
timeout -k 10 900 python -m pytest tests/test_gpu_model.py -q -m gpu -p no:cacheprovider > gpurun_out/model.log 2>&1
echo "model exit=$?" | tee -a gpurun_out/summary.txt; tail -15 gpurun_out/model.log
timeout -k 10 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -p no:cacheprovider -k umma_probe > gpurun_out/probe.log 2>&1
echo "probe exit=$?" | tee -a gpurun_out/summary.txt; tail -3 gpurun_out/probe.log
timeout -k 10 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit=$?" | tee -a gpurun_out/summary.txt; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ "$1" != "noprof" ]; then
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_launch.log 2>&1
echo "ncu-launches exit=$?" | tee -a gpurun_out/summary.txt
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:conv_fwd_umma_kernel -s 12 -c 4 -o gpurun_out/prof_convfwd \
   python bench.py --steps 1 --warmup 3 --no-extras > gpurun_out/ncu_full.log 2>&1
echo "ncu-full exit=$?" | tee -a gpurun_out/summary.txt
fi
cat gpurun_out/summary.txt
