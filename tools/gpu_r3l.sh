#!/bin/bash
# ncu --set full capture of one whole cfg2 VQT pass (8 level kernels + edge frames), final build
mkdir -p gpurun_out
timeout -k 10 500 ncu --set full --clock-control none --import-source on -k regex:"vqt_level_kernel|vqt_edge_kernel" -s 18 -c 9 -f -o gpurun_out/prof_vqt_final python tools/vqt_bench.py 1 > gpurun_out/r3l_ncu.log 2>&1
echo "ncu exit=$?"; tail -2 gpurun_out/r3l_ncu.log; ls -la gpurun_out/prof_vqt_final.ncu-rep
