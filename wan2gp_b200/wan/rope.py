"""3-axis RoPE tables for Wan (hot-path row W0).

Mirrors models/wan/modules/posemb_layers.py::get_rotary_pos_embed (:492-525) ->
get_nd_rotary_pos_embed (:346-431) -> get_1d_rotary_pos_embed (:434-482): head dim 128 split
[44, 42, 42] over the (t, h', w') patch grid, theta 10000, cos/sin repeat-interleaved by 2, fp32.
Computed once per generation on the host (it is 2 x [L,128] floats), then kept resident on the GPU.
"""
import math

import torch

ROPE_DIMS = (44, 42, 42)


def get_rotary_pos_embed(latents_size, enable_RIFLEx=False, theta=10000.0):
    """latents_size = (T, H, W) of the latent; patch (1,2,2).  Returns (cos, sin), each fp32 [L, 128]."""
    T, H, W = (int(s) for s in latents_size)
    if H % 2 or W % 2:
        raise ValueError(f"latent size {latents_size} not divisible by patch size (1,2,2)")
    sizes = (T, H // 2, W // 2)
    grids = torch.meshgrid(*[torch.arange(n, dtype=torch.float32) for n in sizes], indexing="ij")
    cos, sin = [], []
    for axis, (d, g) in enumerate(zip(ROPE_DIMS, grids)):
        freqs = 1.0 / (theta ** (torch.arange(0, d, 2, dtype=torch.float32)[: d // 2] / d))
        if axis == 0 and enable_RIFLEx:
            # RIFLEx (posemb_layers.py:35-86): the intrinsic temporal frequency k=6 is lowered so that one
            # period spans the test length
            k, L_test = 6, T
            freqs[k - 1] = 0.9 * 2 * math.pi / L_test
        ang = torch.outer(g.reshape(-1), freqs)
        cos.append(ang.cos().repeat_interleave(2, dim=1))
        sin.append(ang.sin().repeat_interleave(2, dim=1))
    return torch.cat(cos, 1), torch.cat(sin, 1)
