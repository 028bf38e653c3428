#!/bin/bash
# ncu --set full captures of the dominant launches: cv4/cv5/cv6 forward (conv_fwd_umma_kernel) and the cv4 weight gradient
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:conv_fwd_umma_kernel -s 23 -c 3 -f -o gpurun_out/prof_convfwd_big \
   python bench.py --steps 1 --warmup 3 --no-extras > gpurun_out/ncu_full_big.log 2>&1
echo "ncu-full-fwd exit=$?"
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:"conv_wgrad_umma_kernel|conv_fwd_stack" -s 27 -c 3 -f -o gpurun_out/prof_wgrad_stack \
   python bench.py --steps 1 --warmup 3 --no-extras > gpurun_out/ncu_full_wgrad.log 2>&1
echo "ncu-full-wgrad exit=$?"
