#!/bin/bash
# parity of the VAE kernels, then clip-size timings of the three decoders (run on a GPU box)
cd "$(dirname "$0")/.."
python -c "from wan2gp_b200 import build; build.build()" >/dev/null 2>&1
timeout 600 python -m pytest tests/test_vae_gpu.py tests/test_hy_gpu.py tests/test_edge_gpu.py -q -m gpu -k "vae or conv or rms or group_norm or attention_1head" 2>&1 | tail -5
timeout 300 python tools/wanvae_bench.py 2>&1 | tail -1
timeout 300 python tools/hyvae_bench.py hyvae10 hyvae15 2>&1 | tail -2
