#!/bin/bash
# round 2 session A: full -m gpu suite (fp16 forward activations), then cfg2 VQT timing of the pyramid vs the legacy path
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -x --deselect tests/test_gpu_model.py::test_multi_gpu_ddp_check_if_available > gpurun_out/r2a_tests.log 2>&1
echo "tests exit=$?" | tee gpurun_out/r2a_summary.txt
tail -30 gpurun_out/r2a_tests.log
timeout -k 10 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r2a_tests_all.log 2>&1
echo "tests(all) exit=$?" | tee -a gpurun_out/r2a_summary.txt
tail -15 gpurun_out/r2a_tests_all.log | cut -c1-200
(timeout 300 python tools/vqt_bench.py 5; ZNS_VQT_LEGACY=1 timeout 300 python tools/vqt_bench.py 5) 2>&1 | tee gpurun_out/r2a_vqt.txt
