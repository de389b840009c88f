#!/bin/bash
# round 2, GPU call 17 (8 GPUs): BASELINE configs[4] at 8 GPUs -- one clip per rank, T in {9, 33}, 720p / 1080p, all three decoders
mkdir -p gpurun_out; rm -f gpurun_out/vae_sweep_n8.jsonl
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29761 tools/vae_sweep.py --T 9,33 --res 720p,1080p --max-seconds 6 > gpurun_out/call17_sweep.log 2>&1; echo "rc=$?"; grep '^{' gpurun_out/call17_sweep.log | cut -c1-330
tail -3 gpurun_out/call17_sweep.log | cut -c1-300
