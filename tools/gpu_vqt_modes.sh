#!/bin/bash
# knock-out experiments (diagnostic): 1 skip fp32 conversion, 2 skip epilogue work, 4 skip MMAs, 8 skip loader copies
for m in 0 1 2 4 8 9 6 11 13 14 15; do
  echo "mode $m: $(ZNS_VQT_DBG_MODE=$m timeout 120 python tools/vqt_bench.py 3 --timing 2>&1 | grep -E "cfg2|L[0-3]:" | sed 's/(.*audio.*algorithmic)//' | cut -c1-110 | tr '\n' ' ')"
done
