#!/bin/bash
# round 2, GPU call 18 (1 GPU): GEMM N-group rasterisation A/B, attn6 exp2-split variants
mkdir -p gpurun_out
for ng in 16 8 4 27 54 16; do
  echo "== B200_GEMM_NGROUP=$ng =="; B200_GEMM_NGROUP=$ng timeout 300 python tools/kernel_bench.py gemm 2>&1 | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('   ', d.get('name', d)[:40] if isinstance(d.get('name'), str) else d, round(d.get('ms', 0), 3), round(d.get('tflops', 0)))
" | tee -a gpurun_out/gemm_ngroup_r02.txt
done
echo "== attention exp2-split variants =="; ATT_AB_OUT=attn_ab_call18.json timeout 1200 python tools/attn_ab.py v614 v615 v616 v610 v613 v614 > gpurun_out/call18_attn.log 2>&1; echo "rc=$?"; python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/attn_ab_call18.json"))
    for k, v in d.items():
        print(k, [(t["L"], round(t["ms"], 2), round(t["tflops"])) for t in v.get("timing", [])], all(p["ok"] for p in v.get("parity", [])), v.get("error", "")[-300:])
except Exception as e:
    print("parse failed", e)
PY
