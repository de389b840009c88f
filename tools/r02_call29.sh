#!/bin/bash
# round 2, GPU call 29 (1 GPU): the cp.async-pipelined q|k RMSNorm + RoPE kernel -- bit-identity test, A/B timing against the row kernel, full suite with it as the default
mkdir -p gpurun_out
echo "== rmsnorm tests =="; timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "rmsnorm" 2>&1 | tail -4
echo "== A/B =="; for v in 0 1; do B200_RMSROPE_PIPE=$v timeout 200 python tools/kernel_bench.py rows 2>&1 | grep -E "rmsnorm" ; done | tee gpurun_out/rmsrope_pipe_ab_r02.txt | cut -c1-260
echo "== full GPU suite (pipelined kernel = default) =="; timeout 900 python -m pytest tests -m gpu -q > gpurun_out/call29_tests.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/call29_tests.log
