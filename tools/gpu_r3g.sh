#!/bin/bash
# engine pool (several outstanding forwards per module): model-level tests
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_configs.py -q -m gpu -p no:cacheprovider > gpurun_out/r3g_tests.log 2>&1
echo "tests exit=$?"; tail -5 gpurun_out/r3g_tests.log | cut -c1-300
