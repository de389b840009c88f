#!/bin/bash
# round 2, GPU call 24 (1 GPU): compute-sanitizer memcheck over the umT5 kernels, pipeline / graph tests after the weights-version key
mkdir -p gpurun_out
echo "== pipeline / model / plugin tests =="; timeout 900 python -m pytest tests/test_pipeline_gpu.py tests/test_wan_gpu.py tests/test_plugin_gpu.py tests/test_t5_gpu.py -q -x -m gpu > gpurun_out/call24_tests.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/call24_tests.log
