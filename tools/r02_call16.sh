#!/bin/bash
# round 2, GPU call 16 (1 GPU): the 480p configuration of the 14B model, HunyuanVideo 1.0 720p x 129f
mkdir -p gpurun_out
summ() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "model_tflops", "gpu_launches")}, "attn", d["roofline"]["achieved"], d["roofline"].get("avg_launch_ms"), d["roofline"]["frac"], "vae", d.get("vae_decode", {}).get("ms_per_clip"), d.get("vae_decode", {}).get("value"), d["clocks"], d.get("parity", {}).get("max_rel_l2"))
except Exception as e:
    print("bench parse failed", e)
PY
}
echo "== bench 14B 480p x 81f =="; timeout 900 python bench.py --workload wan22_t2v_14b_480p81 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_480p.json 2> gpurun_out/bench_r02_480p.err; echo "rc=$?"; summ gpurun_out/bench_r02_480p.json; tail -2 gpurun_out/bench_r02_480p.err
echo "== bench HunyuanVideo 1.0 720p x 129f =="; timeout 900 python bench.py --workload hy10_t2v_720p129 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_r02_hy10.json 2> gpurun_out/bench_r02_hy10.err; echo "rc=$?"; summ gpurun_out/bench_r02_hy10.json; tail -2 gpurun_out/bench_r02_hy10.err
