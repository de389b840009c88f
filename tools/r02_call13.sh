#!/bin/bash
# round 2, GPU call 13 (1 GPU): suite after the relaxed-arrive change, Hunyuan 1.5 bench with attn6, ncu launch list of one timed 14B step
mkdir -p gpurun_out
echo "== full GPU suite =="; timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/call13_tests.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/call13_tests.log
echo "== bench (Hunyuan 1.5, configs[3] shape on 1 GPU) =="; timeout 900 python bench.py --workload hy15_t2v_720p129 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_r02_hy15b.json 2> gpurun_out/bench_r02_hy15b.err; echo "rc=$?"; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_r02_hy15b.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "model_tflops", "gpu_launches")}, d["roofline"]["achieved"], d["roofline"]["frac"], d.get("vae_decode", {}).get("ms_per_clip"), d["clocks"])
except Exception as e:
    print("bench parse failed", e)
PY
tail -2 gpurun_out/bench_r02_hy15b.err
echo "== ncu launch list of one timed 14B step =="; B200_CUDA_PROFILER_RANGE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/step_launches_r02.csv python bench.py --steps 1 --warmup 0 --no-vae --no-cpu-baseline > gpurun_out/call13_ncu_step.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/call13_ncu_step.log | cut -c1-300
python tools/launch_summary.py gpurun_out/step_launches_r02.csv 1 > gpurun_out/launches_r02_step.txt 2>&1; head -30 gpurun_out/launches_r02_step.txt
