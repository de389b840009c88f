#!/bin/bash
# round 2, GPU call 3: single-CTA attention variants (packed softmax, exp2 on the FMA pipe), full GPU test suite, VAE head A/B, ncu captures, bench
mkdir -p gpurun_out
echo "== attention A/B =="; timeout 900 python tools/attn_ab.py v1 v100 v104 v103 v102 > gpurun_out/call3_attn.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/call3_attn.log | cut -c1-400
BEST=$(python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/attn_ab.json"))
    ok = {k: v["timing"][0]["ms"] for k, v in d.items() if k[0] == "v" and v.get("timing") and all(p["ok"] for p in v["parity"])}
    print(min(ok, key=ok.get)[1:] if ok else 1)
except Exception:
    print(1)
PY
)
echo "best single-CTA variant: $BEST"; export B200_ATT_VARIANT=$BEST; echo $BEST > gpurun_out/call3_best_variant.txt
echo "== full gpu test suite =="; timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/call3_tests.log 2>&1; echo "rc=$?"; grep -E "passed|failed|FAILED" gpurun_out/call3_tests.log | tail -8
echo "== vae head A/B =="; for f in 1 0; do B200_VAE_HEAD_STACK=$f timeout 300 python tools/wanvae_bench.py 2>&1 | tail -1 | sed "s/^/head_stack=$f: /"; done | tee gpurun_out/call3_vae_ab.log
echo "== ncu launch list of one decode =="; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/wanvae_launches_r02.csv python tools/wanvae_bench.py > gpurun_out/call3_ncu_vae.log 2>&1; echo "rc=$?"
echo "== ncu full: attention + pair gemm =="; timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -c 1 -o gpurun_out/prof_r02_attn -f python tools/profile_targets.py attn > gpurun_out/call3_ncu_attn.log 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_pair -c 1 -o gpurun_out/prof_r02_gemm -f python tools/profile_targets.py gemm > gpurun_out/call3_ncu_gemm.log 2>&1; echo "rc=$?"
echo "== bench =="; timeout 900 python bench.py --steps 3 --warmup 2 > gpurun_out/bench_r02_b.json 2> gpurun_out/bench_r02_b.err; echo "rc=$?"; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_r02_b.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "model_tflops", "gpu_launches")}, d["roofline"]["achieved"], d["roofline"]["avg_launch_ms"], d["gemm_in_step"]["achieved"], d["vae_decode"]["ms_per_clip"], d["clocks"])
except Exception as e:
    print("bench parse failed", e)
PY
tail -3 gpurun_out/bench_r02_b.err
