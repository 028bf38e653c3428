#!/bin/bash
# round 2: tcgen05 VQT pyramid -- parity tests, then cfg2 timing (new default vs ZNS_VQT_LEGACY=1)
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -q -m gpu -p no:cacheprovider -k "vqt or xqt or smoke or step_from_audio" > gpurun_out/vqt2_tests.log 2>&1
echo "vqt tests exit=$?"; tail -15 gpurun_out/vqt2_tests.log
timeout 300 python tools/vqt_bench.py 5
ZNS_VQT_LEGACY=1 timeout 300 python tools/vqt_bench.py 5
