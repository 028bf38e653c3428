#!/bin/bash
mkdir -p gpurun_out
L=$PWD/zeronotesamba_b200
for m in 6 2 4 0; do echo "== groups=$m"; ZNS_VQT_GROUPS=$m timeout 300 python tools/vqt_diff.py $L/libzns_sm100_prev.so 2>&1 | grep -v "clips:\|distinct\|^ \[\|^  *[0-9]" | tail -12; done | tee gpurun_out/r3b_diff.txt
