#!/bin/bash
# A/B of VQT library builds: parity tests on the default build, then cfg2 timing of every zeronotesamba_b200/libzns_*.so
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_configs.py -q -m gpu -p no:cacheprovider -k "vqt or xqt or cfg2 or cfg1" 2>&1 | tail -6
for f in zeronotesamba_b200/libzns_*.so; do
  echo "== $f: $(ZNS_LIB_PATH=$PWD/$f timeout 120 python tools/vqt_bench.py 5 2>&1 | tail -1)"
done | tee gpurun_out/vqt_ab.txt
