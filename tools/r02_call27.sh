#!/bin/bash
# round 2, GPU call 27 (1 GPU): after the fix of the TMA reduce-add epilogue's N tail (found by the byT5 widths in call 26) -- the failing cases again + the new GEMM test
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_llm_gpu.py tests/test_t5_gpu.py tests/test_ops_gpu.py -m gpu -q -s -rA -k "kernels or byt5 or gemm" > gpurun_out/call27_tests.log 2>&1; echo "rc=$?"
grep -E "passed|failed|error" gpurun_out/call27_tests.log | tail -5
grep -E "^(FAILED|ERROR)|^E  " gpurun_out/call27_tests.log | head -40
