#!/bin/bash
# round 2, GPU call 20 (1 GPU): full GPU suite at the committed defaults, smoke, bench line, ncu capture of the conv_row kernel
mkdir -p gpurun_out
echo "== full GPU suite =="; timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/call20_tests.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/call20_tests.log
echo "== smoke =="; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
echo "== bench =="; timeout 900 python bench.py --steps 4 --warmup 3 > gpurun_out/bench_r02_f.json 2> gpurun_out/bench_r02_f.err; echo "rc=$?"; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_r02_f.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "model_tflops", "gpu_launches")}, "attn", d["roofline"]["achieved"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], "gemm", d["gemm_in_step"]["achieved"], "vae", d["vae_decode"]["ms_per_clip"], d["vae_decode"]["value"], d["clocks"], d["parity"].get("max_rel_l2"), d["e2e"]["value"], d.get("cpu_baseline", {}))
except Exception as e:
    print("bench parse failed", e)
PY
tail -3 gpurun_out/bench_r02_f.err
echo "== reference arm =="; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r02_ref2.json 2> gpurun_out/bench_r02_ref2.err; echo "rc=$?"; cut -c1-600 gpurun_out/bench_r02_ref2.json
