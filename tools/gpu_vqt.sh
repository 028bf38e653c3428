#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -q -m gpu -p no:cacheprovider -k "vqt or xqt or smoke or step_from_audio" > gpurun_out/vqt_tests.log 2>&1
echo "vqt tests exit=$?"; tail -5 gpurun_out/vqt_tests.log
ZNS_VQT_UMMA=1 timeout -k 10 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -p no:cacheprovider -k "vqt" > gpurun_out/vqt_tests_umma.log 2>&1
echo "vqt tests (tcgen05 filterbank) exit=$?"; tail -2 gpurun_out/vqt_tests_umma.log
python tools/vqt_bench.py 5
ZNS_VQT_UMMA=1 python tools/vqt_bench.py 3
timeout -k 10 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 30 -c 15 --csv --log-file gpurun_out/vqt_launches.csv python tools/vqt_bench.py 1 > /dev/null 2>&1
grep -v "^==" gpurun_out/vqt_launches.csv | python -c "
import csv,sys,collections
r=csv.DictReader(sys.stdin); agg=collections.OrderedDict()
for row in r:
    k=(row['ID'],row['Kernel Name'][:40]); agg.setdefault(k,{})[row['Metric Name']]=(row['Metric Value'],row['Metric Unit'])
for k,v in agg.items(): print(k, v)
"
