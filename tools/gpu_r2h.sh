#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 1200 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r2h_tests.log 2>&1
echo "tests exit=$?"; tail -8 gpurun_out/r2h_tests.log | cut -c1-220
grep -E "^(tests/.*Error|E   )" gpurun_out/r2h_tests.log | head -20 | cut -c1-220
timeout 900 python bench.py > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench exit=$?"; tail -3 gpurun_out/r2h_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2h_bench.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step','clocks')})
    print('e2e',d['e2e']); print('sustained',{k:v for k,v in d.get('sustained',{}).items() if k!='note'})
    for k,v in d['roofline_kernels'].items(): print('  ',k,{a:round(b,3) for a,b in v.items()})
    print('vqt',d['vqt_cfg2']['ms'],d['vqt_cfg2']['roofline']['frac'])
    print('cfg5',d.get('cfg5_downstream'))
except Exception as e: print('bench parse failed',e)
PY
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2h_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --sustained-s 0 > gpurun_out/r2h_ncu_bench.log 2>&1
python tools/ncu_launch_shares.py gpurun_out/r2h_launches.csv "bench.py --steps 2 --warmup 3 --no-extras under ncu (all launches of the process: 3 warm-up + capture passes + eager count pass; graph replays are not listed)" | head -32
