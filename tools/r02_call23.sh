#!/bin/bash
# round 2, GPU call 23 (1 GPU): compute-sanitizer memcheck over the umT5 kernels, pipeline / graph tests after the weights-version key
mkdir -p gpurun_out
echo "== memcheck: T5 =="; timeout 900 compute-sanitizer --tool memcheck --launch-timeout 300 --error-exitcode 9 python -m pytest tests/test_t5_gpu.py -q -x -m gpu -k "small or wrapper" > gpurun_out/sanitizer_t5_r02.log 2>&1; echo "rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid" gpurun_out/sanitizer_t5_r02.log | head -6
echo "== pipeline / model / plugin tests =="; timeout 900 python -m pytest tests/test_pipeline_gpu.py tests/test_wan_gpu.py tests/test_plugin_gpu.py tests/test_t5_gpu.py -q -x -m gpu > gpurun_out/call23_tests.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/call23_tests.log
