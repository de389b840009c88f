"""Shared helpers for the parity tests (the oracle is imported ONLY from tests)."""
import os

import numpy as np
import torch

from wan2gp_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# name -> (config name, latent (T,H,W), seed)   -- must match oracle/gen_golden.py
WAN_CASES = {
    "tiny": ("tiny", (3, 8, 12), 0),
    "tiny_i2v": ("tiny_i2v", (2, 8, 8), 1),
    "small": ("small", (5, 16, 24), 2),
    "p13b": ("t2v_1.3B", (9, 30, 52), 0),
}
VAE_CASES = {
    "vae_tiny": (synth.VAE_CFG_TINY, (16, 3, 6, 8), 0),
    "vae_small": (synth.VAE_CFG, (16, 5, 8, 10), 1),
    "vae_p": (synth.VAE_CFG, (16, 3, 30, 52), 0),
}


def rel_l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def psnr(a, b, peak):
    mse = float((a.double() - b.double()).pow(2).mean())
    return float("inf") if mse == 0 else 10.0 * float(np.log10(peak * peak / mse))


def load_golden(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: torch.from_numpy(np.asarray(g[k]).astype(np.float32)) if g[k].dtype.kind == "f" else g[k] for k in g.files}


def wan_case(name):
    cfg_name, thw, seed = WAN_CASES[name]
    cfg = synth.WAN_CONFIGS[cfg_name]
    sd = synth.make_wan_state_dict(cfg, seed)
    x, t, ctx, y = synth.make_wan_inputs(cfg, thw, seed)
    return cfg, thw, sd, x, t, ctx, y


def vae_case(name):
    cfg, zshape, seed = VAE_CASES[name]
    sd = synth.make_vae_state_dict(cfg, seed)
    z = synth._normal((1,) + zshape, 1.0, seed, "input.z", "cpu")[0]
    return cfg, sd, z
