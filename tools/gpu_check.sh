#!/bin/bash
# One GPU session: run the -m gpu test groups under individual timeouts, keep every log.
# usage (under gpurun): bash tools/gpu_check.sh [extra pytest -k expression]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
run() { # name, timeout, pytest args...
  local name=$1; shift; local to=$1; shift
  timeout -k 10 "$to" python -m pytest "$@" -q -m gpu -p no:cacheprovider > "gpurun_out/$name.log" 2>&1
  echo "$name exit=$?" | tee -a gpurun_out/summary.txt
  tail -5 "gpurun_out/$name.log"
}
: > gpurun_out/summary.txt
run probe 300 tests/test_gpu_ops.py -k "umma_probe"
run simple 600 tests/test_gpu_ops.py -k "not umma and not conv_fwd and not conv_wgrad and not conv_full and not conv_dgrad"
run convfwd 600 tests/test_gpu_ops.py -k "conv_fwd_umma or two_branches or dgrad"
run wgrad 600 tests/test_gpu_ops.py -k "conv_wgrad_umma"
run full 900 tests/test_gpu_ops.py -k "conv_full"
export ZNS_CONV_TRANSPOSED=1
run convfwd_transposed 600 tests/test_gpu_ops.py -k "conv_fwd_umma or two_branches or dgrad or conv_full"
unset ZNS_CONV_TRANSPOSED
export ZNS_CONV_NO_STACK=1
run convfwd_nostack 600 tests/test_gpu_ops.py -k "conv_fwd_umma or two_branches or dgrad or conv_full"
unset ZNS_CONV_NO_STACK
export ZNS_CONV_PAIR=0
run conv_single_cta 600 tests/test_gpu_ops.py -k "conv_fwd_umma or two_branches or dgrad or conv_full or conv_wgrad_umma"
unset ZNS_CONV_PAIR
export ZNS_WGRAD_STACK=1
run wgrad_stacked_dy 600 tests/test_gpu_ops.py -k "conv_wgrad_umma or conv_full"
unset ZNS_WGRAD_STACK
cat gpurun_out/summary.txt
