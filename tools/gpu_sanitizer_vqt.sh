#!/bin/bash
# compute-sanitizer over the VQT tests and the cv1 kernels after the round-2 rewrites (programmatic dependent launch, unified
# level-0 loader, sixteen / grouped epilogue warps, single bulk-copy loader warp, new edge-frame kernel, two-frame cv1 forward)
mkdir -p gpurun_out
SEL='vqt_vs_oracle or vqt_silence or vqt_ragged or conv1_fwd_wgrad or conv1_dropout'
for tool in memcheck synccheck racecheck; do
  [ "$tool" = racecheck ] && SELT='vqt_silence or conv1_fwd_wgrad' || SELT="$SEL"
  timeout -k 10 ${SAN_TIMEOUT:-600} compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 9 \
      python -m pytest tests/test_gpu_ops.py -q -m gpu -p no:cacheprovider -x -k "$SELT" > gpurun_out/sanitizer_vqt_$tool.log 2>&1
  rc=$?
  {
    echo "== compute-sanitizer --tool $tool  (exit $rc; 9 = errors reported, 124 = timeout)"
    echo "   tests: -k \"$SELT\""
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Error:|Hazard|=========     at " gpurun_out/sanitizer_vqt_$tool.log | sort | uniq -c | sort -rn | head -20
  } | tee gpurun_out/sanitizer_vqt_$tool.txt
done
