#!/bin/bash
# knock-out experiments on the timing build (diagnostic): 1 skip fp32 split, 2 skip epilogue work, 4 skip MMAs, 8 skip loader copies,
# 16 skip epilogue global stores
export ZNS_LIB_PATH=$PWD/zeronotesamba_b200/libzns_sm100_timing.so
mkdir -p gpurun_out
for m in 0 1 2 16 4 8 6 10 12 14; do
  echo "== mode $m"
  ZNS_VQT_DBG_MODE=$m timeout 120 python tools/vqt_bench.py 3 --timing 2>&1 | grep -E "cfg2|L[0-7]:" | sed 's/(.*audio.*algorithmic)//' | awk '{ if ($1 ~ /^L/) { print $1, $2, $6, "|", $(NF-14), $(NF-13), $(NF-12), $(NF-11), $(NF-10), $(NF-9) } else print }' | cut -c1-150
done 2>&1 | tee gpurun_out/vqt_modes.txt
