#!/bin/bash
# round 2, GPU call 21 (1 GPU): whole-step CUDA graph -- bit-identity test, pipeline / model tests, 1.3B bench with and without it
mkdir -p gpurun_out
echo "== tests =="; timeout 900 python -m pytest tests/test_pipeline_gpu.py tests/test_wan_gpu.py tests/test_plugin_gpu.py -q -x -m gpu > gpurun_out/call21_tests.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/call21_tests.log
for v in 1 0; do
  echo "== bench 1.3B, B200_STEP_GRAPH=$v =="; B200_STEP_GRAPH=$v timeout 600 python bench.py --workload wan21_t2v_1.3b_p --steps 20 --warmup 4 --no-cpu-baseline > gpurun_out/bench_r02_13b_sg$v.json 2> gpurun_out/bench_r02_13b_sg$v.err; echo "rc=$?"; python - "gpurun_out/bench_r02_13b_sg$v.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "model_tflops", "model_tensor_frac", "gpu_launches", "finite")}, d["e2e"]["value"], d["config"].get("cuda_graph"), d.get("parity", {}).get("max_rel_l2"))
except Exception as e:
    print("bench parse failed", e)
PY
  tail -2 gpurun_out/bench_r02_13b_sg$v.err
done
