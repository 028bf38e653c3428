#!/bin/bash
# Round-end evidence run on one B200: all GPU tests, smoke, both bench arms, ncu launch list, DRAM traffic of
# the tensor-core kernels, one full capture of the dominant kernel, VQT launch list.
mkdir -p gpurun_out
bash tools/gpu_check.sh > /dev/null 2>&1; cp gpurun_out/summary.txt gpurun_out/summary_ops.txt
timeout -k 10 900 python -m pytest tests/test_gpu_model.py -q -m gpu -p no:cacheprovider > gpurun_out/model.log 2>&1
echo "model exit=$?" >> gpurun_out/summary_ops.txt
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit=$?" >> gpurun_out/summary_ops.txt
timeout -k 10 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit=$?" >> gpurun_out/summary_ops.txt
timeout -k 10 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
echo "bench-reference exit=$?" >> gpurun_out/summary_ops.txt
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 260 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_launch.log 2>&1
echo "ncu-launches exit=$?" >> gpurun_out/summary_ops.txt
timeout -k 10 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:conv_ -s 42 -c 42 --csv \
   --log-file gpurun_out/conv_traffic.csv python bench.py --steps 1 --warmup 3 --no-extras > gpurun_out/ncu_traffic.log 2>&1
echo "ncu-traffic exit=$?" >> gpurun_out/summary_ops.txt
python tools/ncu_traffic_json.py gpurun_out/conv_traffic.csv > gpurun_out/conv_traffic.json 2>/dev/null
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:conv_fwd_umma_kernel -s 16 -c 3 -o gpurun_out/prof_convfwd \
   python bench.py --steps 1 --warmup 3 --no-extras > gpurun_out/ncu_full.log 2>&1
echo "ncu-full exit=$?" >> gpurun_out/summary_ops.txt
python tools/vqt_bench.py 5 > gpurun_out/vqt_bench.txt 2>&1
timeout -k 10 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 30 -c 15 --csv --log-file gpurun_out/vqt_launches.csv python tools/vqt_bench.py 1 > /dev/null 2>&1
cat gpurun_out/summary_ops.txt; cat gpurun_out/vqt_bench.txt; tail -2 gpurun_out/smoke.log
