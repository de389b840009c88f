#!/bin/bash
# round 2, GPU call 22 (1 GPU): the committed state once more -- full GPU suite, smoke, 1.3B bench (launch accounting with the whole-step graph), default bench line
mkdir -p gpurun_out
echo "== full GPU suite =="; timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/call22_tests.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/call22_tests.log
echo "== smoke =="; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
summ() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "model_tflops", "gpu_launches", "finite")}, "attn", d["roofline"]["achieved"], d["roofline"]["frac"], "vae", d.get("vae_decode", {}).get("ms_per_clip"), d["clocks"], d.get("parity", {}).get("max_rel_l2"), d["e2e"]["value"])
except Exception as e:
    print("bench parse failed", e)
PY
}
echo "== bench 1.3B =="; timeout 600 python bench.py --workload wan21_t2v_1.3b_p --steps 20 --warmup 4 --no-cpu-baseline > gpurun_out/bench_r02_13b_final.json 2> gpurun_out/bench_r02_13b_final.err; echo "rc=$?"; summ gpurun_out/bench_r02_13b_final.json
echo "== bench default =="; timeout 900 python bench.py > gpurun_out/bench_r02_g.json 2> gpurun_out/bench_r02_g.err; echo "rc=$?"; summ gpurun_out/bench_r02_g.json; tail -2 gpurun_out/bench_r02_g.err
