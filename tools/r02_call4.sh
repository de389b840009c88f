#!/bin/bash
# round 2, GPU call 4: 16-softmax-warp attention, pair conv, stacked CFG streams (1.3B bench), VAE launch list
mkdir -p gpurun_out
echo "== attention A/B =="; timeout 900 python tools/attn_ab.py v100 v103 v200 v203 > gpurun_out/call4_attn.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/call4_attn.log | cut -c1-120; python - <<'PY'
import json
d = json.load(open("gpurun_out/attn_ab.json"))
for k, v in d.items():
    print(k, v.get("name"), [(t["L"], round(t["ms"], 2), round(t["tflops"])) for t in v.get("timing", [])], v.get("error", "")[:300], [p["ok"] for p in v.get("parity", [])])
PY
BEST=$(python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/attn_ab.json"))
    ok = {k: v["timing"][0]["ms"] for k, v in d.items() if k[0] == "v" and v.get("timing") and all(p["ok"] for p in v["parity"])}
    print(min(ok, key=ok.get)[1:] if ok else 100)
except Exception:
    print(100)
PY
)
echo "best variant: $BEST"; export B200_ATT_VARIANT=$BEST; echo $BEST > gpurun_out/call4_best_variant.txt
echo "== gpu test suite (pair conv, stacked streams, batched attention) =="; timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/call4_tests.log 2>&1; echo "rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/call4_tests.log | tail -8
echo "== conv pair A/B (Wan + Hunyuan VAE decode) =="; for f in 1 0; do B200_CONV_PAIR=$f timeout 300 python tools/wanvae_bench.py 2>&1 | tail -1 | sed "s/^/conv_pair=$f: /"; done | tee gpurun_out/call4_vae_ab.log
for f in 1 0; do B200_CONV_PAIR=$f timeout 400 python tools/hyvae_bench.py hyvae10 hyvae15 2>&1 | tail -2 | sed "s/^/conv_pair=$f: /"; done | tee -a gpurun_out/call4_vae_ab.log
echo "== launch list of one Wan decode =="; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"conv_row|gemm_tcgen05|head_gather|rms_silu|softmax_rows|frames_to|vae_prologue" -c 400 --csv --log-file gpurun_out/wanvae_launches_r02.csv python tools/wanvae_bench.py > gpurun_out/call4_ncu_vae.log 2>&1; echo "rc=$?"
echo "== 1.3B bench (stacked CFG streams + per-block graphs) =="; timeout 600 python bench.py --workload wan21_t2v_1.3b_p --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_13b.json 2> gpurun_out/bench_r02_13b.err; echo "rc=$?"; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_r02_13b.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "model_tflops", "model_tensor_frac", "gpu_launches")}, d.get("parity"))
except Exception as e:
    print("bench parse failed", e)
PY
tail -3 gpurun_out/bench_r02_13b.err
