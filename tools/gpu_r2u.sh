#!/bin/bash
# f1: length-bucketed validation test; ncu --set full of the SIMT kernels next to the conv path (cv1 forward / weight gradient, bias gradient)
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests -q -m gpu -p no:cacheprovider -k "length_buckets or finetune or batched_inference" > gpurun_out/r2u_tests.log 2>&1
echo "tests exit=$?"; tail -3 gpurun_out/r2u_tests.log | cut -c1-300
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:"conv1_fwd_kernel|conv1_wgrad_kernel" -s 4 -c 2 -f -o gpurun_out/prof_cv1 \
   python bench.py --steps 1 --warmup 3 --no-extras --sustained-s 0 > gpurun_out/r2u_ncu.log 2>&1
echo "ncu exit=$?"; tail -2 gpurun_out/r2u_ncu.log
