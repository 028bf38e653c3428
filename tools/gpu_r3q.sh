#!/bin/bash
# final verification after the CQT change: every GPU test, smoke(), memcheck / synccheck over the VQT + CQT op tests
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r3q_tests.log 2>&1
echo "tests exit=$?"; tail -3 gpurun_out/r3q_tests.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -2
for tool in memcheck synccheck; do
  timeout -k 10 300 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 9 python -m pytest tests/test_gpu_ops.py -q -m gpu -p no:cacheprovider -x -k "vqt_vs_oracle or vqt_silence or vqt_ragged" > gpurun_out/sanitizer_cqt_$tool.log 2>&1
  echo "== compute-sanitizer --tool $tool (exit $?)"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer_cqt_$tool.log | tail -2
done
