#!/bin/bash
# round 2, GPU call 5: K/V multicast attention variants, full bench line, Hunyuan 1.5 bench, VAE sweep on 1 GPU
mkdir -p gpurun_out
echo "== attention A/B =="; timeout 900 python tools/attn_ab.py v103 v300 v303 > gpurun_out/call5_attn.log 2>&1; echo "rc=$?"; python - <<'PY'
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
    print(min(ok, key=ok.get)[1:] if ok else 103)
except Exception:
    print(103)
PY
)
echo "best variant: $BEST"; export B200_ATT_VARIANT=$BEST; echo $BEST > gpurun_out/call5_best_variant.txt
echo "== attention tests with the chosen variant =="; timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_wan_gpu.py "tests/test_prod_shapes_gpu.py::test_attention_production_length" -q -x > gpurun_out/call5_tests.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/call5_tests.log
echo "== bench (Wan2.2 14B) =="; timeout 900 python bench.py --steps 4 --warmup 3 > gpurun_out/bench_r02_c.json 2> gpurun_out/bench_r02_c.err; echo "rc=$?"; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_r02_c.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "model_tflops", "gpu_launches")}, "attn", d["roofline"]["achieved"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], "gemm", d["gemm_in_step"]["achieved"], "vae", d["vae_decode"]["ms_per_clip"], d["vae_decode"]["value"], d["clocks"], d["parity"].get("max_rel_l2"), d["e2e"]["value"], d.get("cpu_baseline", {}).get("value"))
except Exception as e:
    print("bench parse failed", e)
PY
tail -3 gpurun_out/bench_r02_c.err
echo "== bench (Hunyuan 1.5, configs[3] shape on 1 GPU) =="; timeout 900 python bench.py --workload hy15_t2v_720p129 --steps 2 --warmup 1 > gpurun_out/bench_r02_hy15.json 2> gpurun_out/bench_r02_hy15.err; echo "rc=$?"; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_r02_hy15.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "model_tflops", "gpu_launches")}, d["roofline"]["achieved"], d.get("vae_decode", {}).get("ms_per_clip"))
except Exception as e:
    print("bench parse failed", e)
PY
tail -3 gpurun_out/bench_r02_hy15.err
echo "== VAE sweep, 1 GPU =="; rm -f gpurun_out/vae_sweep_n1.jsonl; timeout 1200 python tools/vae_sweep.py --max-seconds 25 > gpurun_out/call5_sweep.log 2>&1; echo "rc=$?"; cut -c1-330 gpurun_out/call5_sweep.log | tail -26
