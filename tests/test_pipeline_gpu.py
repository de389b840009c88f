"""GPU: the sampling-loop call sites (schedule, expert switch, CFG pair, fused combine + Euler update, VAE decode to uint8)
against the oracle loop built from the same pieces."""
import pytest
import torch

from tests.helpers import rel_l2, vae_case, wan_case
from wan2gp_b200 import synth

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


def test_denoise_loop_and_decode_match_oracle():
    from oracle import vae_oracle, wan_oracle
    from wan2gp_b200.pipeline import WanDenoiser, euler_timesteps
    from wan2gp_b200.wan import WanModel, WanVAE
    cfg, thw, sd, x, t, ctx, _ = wan_case("tiny")
    sd2 = synth.make_wan_state_dict(cfg, seed=5)                       # "low-noise expert"
    m1, m2 = WanModel(**cfg), WanModel(**cfg)
    m1.load_state_dict(sd), m2.load_state_dict(sd2)
    vcfg, vsd, _ = vae_case("vae_tiny")
    vae = WanVAE(device="cuda", state_dict=vsd, cfg=vcfg)
    steps, shift, g1, g2, thr = 4, 5.0, 4.0, 3.0, 600
    den = WanDenoiser(m1, m2, vae=vae, num_steps=steps, shift=shift, guide_scale=g1, guide2_scale=g2, switch_threshold=thr)
    ctx_null = torch.zeros_like(ctx)
    res = den.generate(ctx, ctx_null, (16,) + thw, seed=11)
    assert res["x"].dtype == torch.uint8 and res["x"].shape == (3, 4 * (thw[0] - 1) + 1, 8 * thw[1], 8 * thw[2])
    # oracle loop
    ts = euler_timesteps(steps, shift)
    lat = torch.randn(1, 16, *thw, generator=torch.Generator().manual_seed(11))
    used = set()
    for i in range(steps):
        tt = torch.tensor([ts[i]])
        w, gs = (sd2, g2) if ts[i] <= thr else (sd, g1)
        used.add(gs)
        c = wan_oracle.wan_forward(w, cfg, lat, tt, ctx, emulate_bf16=True)
        u = wan_oracle.wan_forward(w, cfg, lat, tt, ctx_null, emulate_bf16=True)
        lat = wan_oracle.euler_step(lat, wan_oracle.cfg_combine(c, u, gs), ts[i] / 1000.0, ts[i + 1] / 1000.0)
    assert used == {g1, g2}                                            # both experts were exercised
    got = den.generate(ctx, ctx_null, (16,) + thw, seed=11, decode=False)["latents"].cpu()
    r = rel_l2(got, lat)
    print(f"4-step denoise loop: latents vs oracle loop rel-L2 {r:.3e}")
    assert r < 1e-2
    ref8 = vae_oracle.frames_to_uint8(vae_oracle.vae_decode(vsd, lat[0], synth.VAE_MEAN, synth.VAE_STD, vcfg, emulate_bf16=True))
    d = (res["x"].int() - ref8.int()).abs().float()
    print(f"decoded frames vs oracle: mean |d uint8| {d.mean():.3f}")
    assert d.mean() < 2.0

    # interrupt from the "UI thread" aborts the whole generation with None (any2video.py:1995-1998 -> wgp.py:7937)
    calls = []

    def cb(step, latent, force, read_state):
        calls.append(1)
        if len(calls) == 3:
            den._interrupt = True
    assert den.generate(ctx, ctx_null, (16,) + thw, seed=11, callback=cb) is None
