#!/bin/bash
# round 2, GPU call 1: pair GEMM parity + A/B timing, library bar, production-shape parity tests
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/call1_smi.txt 2>&1
echo "== pair gemm ==" ; timeout 900 python tools/pair_gemm_check.py time > gpurun_out/call1_pair.log 2>&1; echo "rc=$?"; tail -40 gpurun_out/call1_pair.log
echo "== ops tests ==" ; timeout 900 python -m pytest tests/test_ops_gpu.py -x -q > gpurun_out/call1_ops.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/call1_ops.log
echo "== lib bar ==" ; timeout 900 python tools/kernel_bench.py lib libconv rows > gpurun_out/call1_lib.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/call1_lib.log
echo "== prod shapes ==" ; timeout 1500 python -m pytest tests/test_prod_shapes_gpu.py -q -s > gpurun_out/call1_prod.log 2>&1; echo "rc=$?"; tail -40 gpurun_out/call1_prod.log
