#!/bin/bash
# compute-sanitizer over the op-level GPU tests of every hand-written kernel family:
# memcheck, synccheck, racecheck, initcheck.  Summaries land in gpurun_out/sanitizer_*.txt (copied to profiles/ by hand).
mkdir -p gpurun_out
SEL='conv_fwd_umma or two_branches or dgrad or conv_wgrad_umma or vqt_vs_oracle or vqt_silence or ntxent or pool_fwd_bwd or head_fwd_bwd or conv1_fwd_wgrad or bias_grad or pack_unpack or crop_gather'
for tool in memcheck synccheck racecheck; do
  extra=""
  [ "$tool" = racecheck ] && SELT='conv_fwd_umma or conv_wgrad_umma or ntxent or vqt_silence' || SELT="$SEL"
  timeout -k 10 ${SAN_TIMEOUT:-900} compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 9 $extra \
      python -m pytest tests/test_gpu_ops.py -q -m gpu -p no:cacheprovider -x -k "$SELT" > gpurun_out/sanitizer_$tool.log 2>&1
  rc=$?
  {
    echo "== compute-sanitizer --tool $tool  (exit $rc; 9 = errors reported, 124 = timeout)"
    echo "   tests: -k \"$SELT\""
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Error:|Hazard|=========     at " gpurun_out/sanitizer_$tool.log | sort | uniq -c | sort -rn | head -20
  } | tee gpurun_out/sanitizer_$tool.txt
done
