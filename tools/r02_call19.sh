#!/bin/bash
# round 2, GPU call 19 (1 GPU): umT5 text encoder -- parity tests, timing at the XXL size
mkdir -p gpurun_out
echo "== T5 tests =="; timeout 900 python -m pytest tests/test_t5_gpu.py -q -x -m gpu > gpurun_out/call19_tests.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/call19_tests.log
echo "== umT5-XXL timing =="; true
