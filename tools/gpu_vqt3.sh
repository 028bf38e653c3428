#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/vqt_diag.py 30
ZNS_VQT_LEGACY=1 timeout 300 python tools/vqt_diag.py 30
timeout -k 10 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 18 -c 9 --csv --log-file gpurun_out/vqt2_launches.csv python tools/vqt_bench.py 1 > /dev/null 2>&1
grep -v "^==" gpurun_out/vqt2_launches.csv | python -c "
import csv,sys,collections
r=csv.DictReader(sys.stdin); agg=collections.OrderedDict()
for row in r:
    k=(row['ID'],row['Kernel Name'][:40],row['Grid Size'],row['Block Size']); agg.setdefault(k,{})[row['Metric Name']]=row['Metric Value']
for k,v in agg.items(): print(k, v)
"
