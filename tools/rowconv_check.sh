#!/bin/bash
# A/B check of the row-tiled conv kernel on a GPU box: VAE parity suites with the default selection and with the row
# kernel forced for every spatial conv, then clip-size timings of the three decoders with and without it.
cd "$(dirname "$0")/.."
python -c "from wan2gp_b200 import build; build.build()" >/dev/null 2>&1
for mode in 1 2; do
  echo "== B200_CONV_ROW=$mode"
  B200_CONV_ROW=$mode timeout 600 python -m pytest tests/test_vae_gpu.py tests/test_hy_gpu.py -x -q -m gpu -k "vae or conv" 2>&1 | tail -4
done
for mode in 1 0; do
  echo "== timings, B200_CONV_ROW=$mode"
  B200_CONV_ROW=$mode timeout 300 python tools/wanvae_bench.py 2>&1 | tail -1
  B200_CONV_ROW=$mode timeout 300 python tools/hyvae_bench.py hyvae10 hyvae15 2>&1 | tail -2
done
