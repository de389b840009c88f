#!/bin/bash
# round 2, GPU call 11 (1 GPU): three-score-buffer attention kernel (attn6) A/B against the default
mkdir -p gpurun_out
echo "== attention A/B =="; ATT_AB_OUT=attn_ab_call11.json timeout 1500 python tools/attn_ab.py v103 v603 v604 v613 v614 v612 v103 > gpurun_out/call11_attn.log 2>&1; echo "rc=$?"; python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/attn_ab_call11.json"))
    for k, v in d.items():
        print(k, v.get("name"), [(t["L"], round(t["ms"], 2), round(t["tflops"])) for t in v.get("timing", [])], [(p["Lq"], p["Lk"], p.get("boost"), "%.2e" % p["rel_l2"]) for p in v.get("parity", [])], v.get("error", "")[-800:])
except Exception as e:
    print("parse failed", e)
PY
tail -5 gpurun_out/call11_attn.log
