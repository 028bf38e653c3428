#!/bin/bash
# evidence run: cfg2 VQT launch list with DRAM traffic, role counters, full GPU tests, default bench line
mkdir -p gpurun_out
timeout -k 10 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 18 -c 9 --csv --log-file gpurun_out/r02_vqt_launches.csv python tools/vqt_bench.py 1 > /dev/null 2>&1
python tools/vqt_bench.py 5 | tee gpurun_out/r02_vqt_bench.txt
ZNS_LIB_PATH=$PWD/zeronotesamba_b200/libzns_sm100_timing.so python tools/vqt_bench.py 3 --timing > gpurun_out/r02_vqt_role_counters.txt 2>&1

timeout -k 10 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r3f_tests.log 2>&1
echo "tests exit=$?"; tail -3 gpurun_out/r3f_tests.log | cut -c1-200
timeout 900 python bench.py > gpurun_out/r3f_bench.json 2> gpurun_out/r3f_bench.err; echo "bench exit=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r3f_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step')}); print('e2e',d['e2e']['value']); print('sus',d['sustained']['value'],d['sustained']['clocks'])
print('vqt',d['vqt_cfg2']['ms'],d['vqt_cfg2']['roofline']['frac'],d['vqt_cfg2']['roofline']['traffic'])"
for tool in memcheck synccheck; do
  timeout -k 10 300 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 9 python -m pytest tests/test_gpu_ops.py -q -m gpu -p no:cacheprovider -x -k "vqt_vs_oracle or vqt_silence or vqt_ragged or bias_grad" > gpurun_out/sanitizer_final_$tool.log 2>&1
  echo "== compute-sanitizer --tool $tool (exit $?)"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer_final_$tool.log | tail -2
done
