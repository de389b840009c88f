#!/bin/bash
# round 2, GPU call 12 (1 GPU): attn6 as the default -- full GPU suite, bench A/B against the two-Q-tile kernel (14B and 1.3B), ncu capture
mkdir -p gpurun_out
echo "== full GPU suite =="; timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/call12_tests.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/call12_tests.log
summ() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "model_tflops", "gpu_launches")}, "attn", d["roofline"]["kernel"][:24], d["roofline"]["achieved"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], "gemm", d.get("gemm_in_step", {}).get("achieved"), "vae", d.get("vae_decode", {}).get("ms_per_clip"), d["clocks"], d.get("parity", {}).get("max_rel_l2"), d["e2e"]["value"])
except Exception as e:
    print("bench parse failed", e)
PY
}
for v in 614 103; do
  echo "== bench 14B, B200_ATT_VARIANT=$v =="; B200_ATT_VARIANT=$v timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_e_$v.json 2> gpurun_out/bench_r02_e_$v.err; echo "rc=$?"; summ gpurun_out/bench_r02_e_$v.json; tail -2 gpurun_out/bench_r02_e_$v.err
done
for v in 614 103; do
  echo "== bench 1.3B, B200_ATT_VARIANT=$v =="; B200_ATT_VARIANT=$v timeout 600 python bench.py --workload wan21_t2v_1.3b_p --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_13b_$v.json 2> gpurun_out/bench_r02_13b_$v.err; echo "rc=$?"; summ gpurun_out/bench_r02_13b_$v.json
done
echo "== ncu attn_s3 =="; timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_s3 -c 1 -o gpurun_out/prof_r02_attn6 -f python tools/profile_targets.py attn > gpurun_out/call12_ncu.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/call12_ncu.log
python tools/ncu_summary.py gpurun_out/prof_r02_attn6.ncu-rep > gpurun_out/ncu_r02_attn6.txt 2>&1; head -12 gpurun_out/ncu_r02_attn6.txt
