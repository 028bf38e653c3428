#!/bin/bash
# round 2 session F: full -m gpu suite, bench (default), VQT cfg2 timing, compute-sanitizer summaries
mkdir -p gpurun_out
timeout -k 10 1200 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r2f_tests.log 2>&1
echo "tests exit=$?"; tail -12 gpurun_out/r2f_tests.log | cut -c1-220
grep -E "^E  " gpurun_out/r2f_tests.log | head -20 | cut -c1-220
python tools/vqt_bench.py 5
timeout 900 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench exit=$?"; tail -3 gpurun_out/r2f_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2f_bench.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step','clocks')})
    print('e2e',d['e2e']); print('sustained',{k:v for k,v in d.get('sustained',{}).items() if k!='note'})
    for k,v in d['roofline_kernels'].items(): print('  ',k,{a:round(b,3) for a,b in v.items()})
    print('vqt',d['vqt_cfg2']['ms'],d['vqt_cfg2']['roofline']['frac'])
except Exception as e: print('bench parse failed',e)
PY
SAN_TIMEOUT=420 bash tools/gpu_sanitizer.sh 2>&1 | tail -40
