#!/bin/bash
# A/B checks of the conv / elementwise kernels on a GPU box: VAE parity suites, then clip-size timings of the three decoders.
cd "$(dirname "$0")/.."
python -c "from wan2gp_b200 import build; build.build()" >/dev/null 2>&1
for k32 in 1 0; do
  echo "== parity, B200_CONV_ROW_K32=$k32"
  B200_CONV_ROW_K32=$k32 timeout 600 python -m pytest tests/test_vae_gpu.py tests/test_hy_gpu.py tests/test_edge_gpu.py -q -m gpu -k "vae or conv or rms or group_norm or attention_1head" 2>&1 | tail -5
done
for k32 in 1 0; do
  echo "== Wan VAE timing, B200_CONV_ROW_K32=$k32"
  B200_CONV_ROW_K32=$k32 timeout 300 python tools/wanvae_bench.py 2>&1 | tail -1
done
timeout 300 python tools/hyvae_bench.py hyvae10 hyvae15 2>&1 | tail -2
