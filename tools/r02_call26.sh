#!/bin/bash
# round 2, GPU call 26 (1 GPU): first GPU run of the new pieces -- Hunyuan boundary levels 1-2 (generate vs the oracle loop), byT5 encoder, LLM text towers
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_llm_gpu.py tests/test_hy_plugin_gpu.py tests/test_t5_gpu.py -m gpu -q -s -rA > gpurun_out/call26_tests.log 2>&1; echo "rc=$?"
grep -E "passed|failed|error" gpurun_out/call26_tests.log | tail -5
grep -E "^(FAILED|ERROR)|rel-L2|vs bf16|hidden_states" gpurun_out/call26_tests.log | head -40
