#!/bin/bash
# round 2, GPU call 8 (1 GPU): TMEM port micro-benchmark, attention limiter ablation, second-generation pair conv kernel (parity, A/B, ncu)
mkdir -p gpurun_out
echo "== TMEM read/write port =="; timeout 120 tools/tmem_bw > gpurun_out/tmem_bw_r02.jsonl 2>&1; echo "rc=$?"; cat gpurun_out/tmem_bw_r02.jsonl | cut -c1-200
echo "== attention limiter ablation =="; ATT_AB_OUT=attn_ablation.json timeout 900 python tools/attn_ab.py v103 v903 v904 v902 v901 v905 v103 > gpurun_out/call8_attn.log 2>&1; echo "rc=$?"; python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/attn_ablation.json"))
    for k, v in d.items():
        print(k, v.get("name"), [(t["L"], round(t["ms"], 2), [round(x, 1) for x in t["all_ms"]]) for t in v.get("timing", [])], v.get("error", "")[:300])
except Exception as e:
    print("parse failed", e)
PY
echo "== conv v2: parity =="; timeout 900 python -m pytest tests/test_vae_gpu.py tests/test_hy_gpu.py tests/test_edge_gpu.py tests/test_prod_shapes_gpu.py -q -m gpu -x -k "vae or conv or decode or encode" > gpurun_out/call8_tests.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/call8_tests.log
for v in 1 0 1; do
  echo "== conv v2=$v: decoders =="
  B200_CONV_V2=$v timeout 300 python tools/wanvae_bench.py 2>&1 | tail -1 | sed "s/^/conv_v2=$v: /" | tee -a gpurun_out/vae_conv_v2_ab.txt
done
for v in 1 0; do
  B200_CONV_V2=$v timeout 400 python tools/hyvae_bench.py hyvae10 hyvae15 2>&1 | tail -2 | sed "s/^/conv_v2=$v: /" | tee -a gpurun_out/vae_conv_v2_ab.txt
done
echo "== ncu conv_row2 =="; timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_row -c 1 -o gpurun_out/ncu_r02_conv2 -f python tools/profile_targets.py conv > gpurun_out/call8_ncu.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/call8_ncu.log
ncu -i gpurun_out/ncu_r02_conv2.ncu-rep --page details > gpurun_out/ncu_r02_conv2.txt 2>&1; grep -E "conv_row|Duration|SM Frequency|Registers Per|TC is" gpurun_out/ncu_r02_conv2.txt | head
ncu -i gpurun_out/ncu_r02_conv2.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h,u,v=rows[0],rows[1],rows[2]
for a,b,c in zip(h,u,v):
    if a in ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed','gpu__time_duration.sum','sm__cycles_elapsed.avg.per_second'): print(a,b,c)
"
