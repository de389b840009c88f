#!/bin/bash
# one ncu --set full capture each of the GroupNorm apply pass and the RMS-norm pass at full resolution
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:group_norm_apply -s 50 -c 1 -o gpurun_out/ncu_gn_apply -f python tools/hyvae_bench.py hyvae10 > gpurun_out/ncu_gn_apply.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rms_silu_cl_kernel -s 30 -c 1 -o gpurun_out/ncu_rms -f python tools/wanvae_bench.py > gpurun_out/ncu_rms.log 2>&1
ls -la gpurun_out/*.ncu-rep
