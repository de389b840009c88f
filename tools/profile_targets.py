"""One launch each of the dominant kernels at BASELINE shapes, for `ncu --set full` captures (see profiles/README.md)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wan2gp_b200 import ops  # noqa: E402
from wan2gp_b200.wan.vae import _Conv  # noqa: E402

bf16, f32 = torch.bfloat16, torch.float32
L, D, F, H = 75600, 5120, 13824, 40
which = sys.argv[1:] or ["attn", "gemm", "conv"]
if "attn" in which:
    qkv = torch.randn(L, 3 * D, device="cuda").to(bf16)
    out = torch.empty(L, D, device="cuda", dtype=bf16)
    ops.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], H, out=out)
    del qkv, out
if "gemm" in which:
    a = torch.randn(L, D, device="cuda").to(bf16)
    w = (torch.randn(F, D, device="cuda") * D ** -0.5).to(bf16)
    ops.gemm(a, w, bias=torch.randn(F, device="cuda"), act=1)
    del a, w
if "conv" in which:
    x = torch.randn(13, 720, 1280, 96, device="cuda").to(bf16)
    conv = _Conv(torch.randn(96, 96, 3, 3, 3, device="cuda") * 0.02, torch.randn(96, device="cuda"), "cuda")
    if conv.fusable(1280):   # the variant the decoder runs: CTA pair + fused next-layer RMS_norm/SiLU
        conv.with_norm(x, torch.ones(96, device="cuda"))
    else:
        conv(x)
if "rows" in which:          # the HBM-bound row kernels of a 14B block: LayerNorm + modulation, q|k RMSNorm + RoPE (one launch)
    x = torch.randn(L, D, device="cuda")
    y = torch.empty(L, D, device="cuda", dtype=bf16)
    ops.ln_modulate(x, torch.randn(D, device="cuda"), torch.randn(D, device="cuda"), out=y)
    del x, y
    qkv = torch.randn(L, 3 * D, device="cuda").to(bf16)
    w = torch.ones(D, device="cuda")
    cos, sin = torch.randn(L, 128, device="cuda"), torch.randn(L, 128, device="cuda")
    ops.qk_rmsnorm_rope_(qkv[:, :D], qkv[:, D:2 * D], w, w, 1e-6, cos, sin)
    del qkv
torch.cuda.synchronize()
print("done")
