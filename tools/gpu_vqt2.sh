#!/bin/bash
# round 2: tcgen05 VQT pyramid -- parity tests, then cfg2 timing (new default vs ZNS_VQT_LEGACY=1), per-kernel times
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -q -m gpu -p no:cacheprovider -k "vqt or xqt or smoke or step_from_audio" > gpurun_out/vqt2_tests.log 2>&1
echo "vqt tests exit=$?"; tail -15 gpurun_out/vqt2_tests.log
timeout 300 python tools/vqt_bench.py 5
ZNS_VQT_LEGACY=1 timeout 300 python tools/vqt_bench.py 5
timeout -k 10 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 18 -c 9 --csv --log-file gpurun_out/vqt2_launches.csv python tools/vqt_bench.py 1 > /dev/null 2>&1
grep -v "^==" gpurun_out/vqt2_launches.csv | python -c "
import csv,sys,collections
r=csv.DictReader(sys.stdin); agg=collections.OrderedDict()
for row in r:
    k=(row['ID'],row['Kernel Name'][:40],row['Grid Size'],row['Block Size']); agg.setdefault(k,{})[row['Metric Name']]=row['Metric Value']
for k,v in agg.items(): print(k, v)
"
