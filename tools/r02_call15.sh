#!/bin/bash
# round 2, GPU call 15 (1 GPU): compute-sanitizer over the new cluster kernels (small shapes), BASELINE configs[4] sweep with the final kernels
mkdir -p gpurun_out
echo "== compute-sanitizer memcheck: attention + conv tests (small shapes) =="
timeout 900 compute-sanitizer --tool memcheck --launch-timeout 300 --error-exitcode 9 python -m pytest tests/test_ops_gpu.py tests/test_vae_gpu.py -q -x -m gpu -k "attention or conv or decode_matches or streamed" > gpurun_out/sanitizer_memcheck_r02.log 2>&1; echo "rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|error" gpurun_out/sanitizer_memcheck_r02.log | head -12
echo "== compute-sanitizer racecheck: attention tests =="
timeout 600 compute-sanitizer --tool racecheck --launch-timeout 300 --error-exitcode 9 python -m pytest tests/test_ops_gpu.py -q -x -m gpu -k "attention" > gpurun_out/sanitizer_racecheck_r02.log 2>&1; echo "rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/sanitizer_racecheck_r02.log | head -12
echo "== VAE sweep, 1 GPU =="; rm -f gpurun_out/vae_sweep_n1.jsonl; timeout 1200 python tools/vae_sweep.py --max-seconds 25 > gpurun_out/call15_sweep.log 2>&1; echo "rc=$?"; cut -c1-300 gpurun_out/call15_sweep.log | tail -26
