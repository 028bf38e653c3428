#!/bin/bash
mkdir -p gpurun_out
for f in zeronotesamba_b200/libzns_nopf.so zeronotesamba_b200/libzns_sm100.so; do
  echo "== $f: $(ZNS_LIB_PATH=$PWD/$f timeout 120 python tools/vqt_bench.py 5 2>&1 | tail -1)"
done | tee gpurun_out/vqt_ab.txt
timeout -k 10 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r2d_tests.log 2>&1
echo "tests exit=$?"; tail -4 gpurun_out/r2d_tests.log | cut -c1-200
timeout 600 python bench.py > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench exit=$?"; tail -3 gpurun_out/r2d_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2d_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step','clocks')})
print('e2e',d['e2e']); print('sustained',d.get('sustained'))
print('roofline',{k:d['roofline'][k] for k in ('kernel','achieved','peak','frac','frac_of_clock_peak','executed_flop_fraction')})
for k,v in d['roofline_kernels'].items(): print('  ',k,{a:round(b,3) for a,b in v.items()})
print('vqt',d['vqt_cfg2']['ms'],d['vqt_cfg2']['roofline']['frac'])
print('cpu',d.get('cpu_baseline'))
PY
