#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_model.py -q -m gpu -p no:cacheprovider -k "fused_adam or train_epoch or trainer" > gpurun_out/r3n_tests.log 2>&1
echo "tests exit=$?"; tail -12 gpurun_out/r3n_tests.log | cut -c1-250
