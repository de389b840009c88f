#!/bin/bash
# round 2, GPU call 9 (1 GPU): TMEM port micro-benchmark (fixed store loop, fragment shapes), double-buffered-score attention kernel A/B, Wan VAE launch list
mkdir -p gpurun_out
echo "== TMEM read/write port =="; timeout 120 tools/tmem_bw > gpurun_out/tmem_bw_r02b.jsonl 2>&1; echo "rc=$?"; cut -c1-170 gpurun_out/tmem_bw_r02b.jsonl
echo "== attention A/B =="; ATT_AB_OUT=attn_ab_call9.json timeout 1200 python tools/attn_ab.py v103 v500 v503 v504 v103 > gpurun_out/call9_attn.log 2>&1; echo "rc=$?"; python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/attn_ab_call9.json"))
    for k, v in d.items():
        print(k, v.get("name"), [(t["L"], round(t["ms"], 2), round(t["tflops"])) for t in v.get("timing", [])], [(p["Lq"], p["Lk"], "%.2e" % p["rel_l2"]) for p in v.get("parity", [])], v.get("error", "")[-600:])
except Exception as e:
    print("parse failed", e)
PY
echo "== launch list of one Wan decode =="; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"conv_row|gemm_tcgen05|head_gather|rms_silu|softmax_rows|frames_to|vae_prologue" -c 400 --csv --log-file gpurun_out/wanvae_launches_r02b.csv python tools/wanvae_bench.py > gpurun_out/call9_ncu_vae.log 2>&1; echo "rc=$?"
python tools/launch_summary.py gpurun_out/wanvae_launches_r02b.csv 4 > gpurun_out/launches_r02b_wanvae.txt 2>&1; head -24 gpurun_out/launches_r02b_wanvae.txt
