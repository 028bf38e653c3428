#!/bin/bash
# final verification of the round: every GPU test, smoke(), both bench arms (reference arm shortened)
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r3m_tests.log 2>&1
echo "tests exit=$?"; tail -3 gpurun_out/r3m_tests.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r3m_bench.json 2> gpurun_out/r3m_bench.err; echo "bench exit=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r3m_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step')}); print('e2e',d['e2e']['value']); print('sus',d['sustained']['value'],d['sustained']['clocks'])
print('vqt',d['vqt_cfg2']['ms'],d['vqt_cfg2']['roofline']['frac']); print('cfg5',d['cfg5_downstream']['inference']['clips_per_sec'],d['cfg5_downstream']['finetune']['files_per_sec'])"
