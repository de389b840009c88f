#!/bin/bash
# round 2, GPU call 28 (1 GPU): the final state -- full GPU suite, smoke, attention vs cuDNN SDPA side by side (default attn6 kernel), ncu capture of
# the HBM-bound row kernels, default-shape bench line (CPU arm skipped here: the driver's own run times it)
mkdir -p gpurun_out
echo "== full GPU suite =="; timeout 900 python -m pytest tests -m gpu -q > gpurun_out/call28_tests.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/call28_tests.log
echo "== smoke =="; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
echo "== libattn =="; timeout 300 python tools/kernel_bench.py libattn > gpurun_out/call28_libattn.log 2>&1; echo "rc=$?"; cat gpurun_out/call28_libattn.log | cut -c1-400
echo "== ncu rows =="; timeout 300 ncu --set full --clock-control none --import-source on -k regex:"ln_modulate|rmsnorm_rope" -c 2 -o gpurun_out/prof_r02_rows -f python tools/profile_targets.py rows > gpurun_out/call28_ncu.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/call28_ncu.log
python tools/ncu_summary.py gpurun_out/prof_r02_rows.ncu-rep > gpurun_out/ncu_r02_rows.txt 2>&1; grep -E "kernel:|gpu__time_duration|dram__bytes|dram_throughput" gpurun_out/ncu_r02_rows.txt | head -12
echo "== llm bench =="; timeout 200 python tools/llm_bench.py > gpurun_out/llm_bench_r02.jsonl 2> gpurun_out/llm_bench.err; echo "rc=$?"; cat gpurun_out/llm_bench_r02.jsonl; tail -2 gpurun_out/llm_bench.err
echo "== bench default shape =="; timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_r02_final.json 2> gpurun_out/bench_r02_final.err; echo "rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_r02_final.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "model_tflops", "gpu_launches", "finite")}, "attn", d["roofline"]["achieved"], d["roofline"]["frac"], "vae", d.get("vae_decode", {}).get("ms_per_clip"), d["clocks"], d.get("parity", {}).get("max_rel_l2"), d["e2e"]["value"])
except Exception as e:
    print("bench parse failed", e)
PY
tail -2 gpurun_out/bench_r02_final.err
