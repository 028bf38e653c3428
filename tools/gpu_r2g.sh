#!/bin/bash
mkdir -p gpurun_out
python tools/diag_varlen.py 2>&1 | tail -20
timeout -k 10 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -p no:cacheprovider -k "conv_pool_fused or two_branch or multi_tensor or bce" 2>&1 | tail -6 | cut -c1-200
# per-launch times of two training steps (graph replay is opaque to ncu: eager step via --no-extras path with use_graph in bench -> steps 2)
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2g_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --sustained-s 0 > gpurun_out/r2g_ncu_bench.log 2>&1
python tools/ncu_launch_shares.py gpurun_out/r2g_launches.csv "bench.py --steps 2 --warmup 3 under ncu (all launches of the process)" | head -40
