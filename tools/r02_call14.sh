#!/bin/bash
# round 2, GPU call 14 (2 GPUs): torchrun product-path test, bench --gpus 2 with the cfg_split (configs[2]) and forced hy15 (configs[3]) sub-runs
mkdir -p gpurun_out

nvidia-smi --query-gpu=index,name --format=csv,noheader
echo "== multi-GPU product-path test =="; timeout 600 python -m pytest tests/test_multigpu_gpu.py tests/test_dist_cpu.py -q -s > gpurun_out/call14_tests.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/call14_tests.log
echo "== bench --gpus 2 (+ cfg_split, + forced hy15 sub-run) =="; B200_BENCH_FORCE_HY15=1 timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29755 bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_r02_n2b.json 2> gpurun_out/bench_r02_n2b.err; echo "rc=$?"; python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_r02_n2b.json").read().strip().splitlines() if l.startswith("{")][-1])
    print({k: d[k] for k in ("value", "ms_per_step", "n_gpus", "gpu_launches")}, d["vae_decode"].get("fused_u8_allgather_ms"), d["vae_decode"].get("u8_plus_nccl_allgather_ms"), d["vae_decode"].get("fused_matches_nccl"))
    print("cfg_split:", {k: d["cfg_split"].get(k) for k in ("value", "ms_per_step", "error")})
    print("hy15:", {k: d["hy15_t2v_720p129"].get(k) for k in ("value", "ms_per_step", "error")})
except Exception as e:
    print("bench parse failed", e)
PY
tail -5 gpurun_out/bench_r02_n2b.err
