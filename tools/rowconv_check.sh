#!/bin/bash
# A/B check of the row-tiled conv kernel on a GPU box: parity with and without the descriptor base offset, then the
# whole VAE suites with the row kernel forced everywhere, then clip-size timings.
cd "$(dirname "$0")/.."
for bo in 1 0; do
  echo "== B200_CONV_ROW_BASEOFF=$bo"
  B200_CONV_ROW_BASEOFF=$bo timeout 300 python -m pytest tests/test_vae_gpu.py -x -q -m gpu -k "row_kernel" 2>&1 | tail -6
done
