#!/bin/bash
# round 2, GPU call 2: attention pair kernel A/B + parity, remaining prod-shape test, plugin test, full bench line
mkdir -p gpurun_out
echo "== attention A/B =="; timeout 1100 python tools/attn_ab.py > gpurun_out/call2_attn.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/call2_attn.log
# pick the default for the rest of this call: the pair kernel only if its parity leg passed
PAIR_OK=$(python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/attn_ab.json")); v = d.get("1", {})
    print(1 if v.get("parity") and all(p["ok"] for p in v["parity"]) and v.get("timing") else 0)
except Exception:
    print(0)
PY
)
echo "pair kernel usable: $PAIR_OK"; export B200_ATT_PAIR=$PAIR_OK
echo "== ops + prod + plugin tests =="; timeout 1200 python -m pytest tests/test_ops_gpu.py tests/test_plugin_gpu.py tests/test_multigpu_gpu.py "tests/test_prod_shapes_gpu.py::test_wanvae_decode_720p_9frames" "tests/test_prod_shapes_gpu.py::test_attention_production_length" tests/test_wan_gpu.py tests/test_hy_gpu.py -q -s -x > gpurun_out/call2_tests.log 2>&1; echo "rc=$?"; grep -E "rel-L2|passed|failed|mean \|d" gpurun_out/call2_tests.log | tail -25
echo "== vae tests + fused-norm A/B =="; timeout 900 python -m pytest tests/test_vae_gpu.py -q -x > gpurun_out/call2_vae.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/call2_vae.log
for f in 1 0; do B200_VAE_FUSE_NORM=$f timeout 300 python tools/wanvae_bench.py 2>&1 | tail -2 | sed "s/^/fuse_norm=$f: /"; done | tee gpurun_out/call2_vae_ab.log
echo "== bench =="; timeout 900 python bench.py --steps 3 --warmup 2 > gpurun_out/bench_r02_a.json 2> gpurun_out/bench_r02_a.err; echo "rc=$?"; tail -c 3000 gpurun_out/bench_r02_a.json; tail -5 gpurun_out/bench_r02_a.err
