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


def test_unipc_fused_step_kernel():
    """b200_cfg_unipc_step over a whole 8-step schedule (CFG pair, CFG-Zero* on from step 2): one fused kernel per step against the
    CPU restatement with the same host coefficients (which tests/test_unipc_cpu.py pins to the reference scheduler)."""
    from oracle import wan_oracle
    from wan2gp_b200 import ops
    from wan2gp_b200.pipeline import UniPCSchedule
    steps, shift, g = 8, 5.0, 4.0
    gen = torch.Generator().manual_seed(2)
    x = torch.randn(1, 16, 3, 8, 10, generator=gen)
    conds = [torch.randn_like(x) for _ in range(steps)]
    unconds = [c * 0.7 + 0.3 * torch.randn(x.shape, generator=gen) for c in conds]
    sch_g, sch_c = UniPCSchedule(steps, shift), UniPCSchedule(steps, shift)
    xg = x.cuda()
    hist = [torch.zeros_like(xg) for _ in range(3)]
    xc_, xl, m0, m1 = x.double(), torch.zeros_like(x).double(), torch.zeros_like(x).double(), torch.zeros_like(x).double()
    for i in range(steps):
        star = i >= 2
        c, u = conds[i].double(), unconds[i].double()
        if star:
            u = u * ((c * u).sum() / ((u * u).sum() + 1e-8))
        v = wan_oracle.cfg_combine(c, u, g)
        xc_, corr, x0 = wan_oracle.unipc_step(xc_, v, xl, m0, m1, sch_c.coefficients(i))
        xl, m0, m1 = corr, x0, m0
        ops.cfg_unipc_step_(xg, conds[i].cuda(), unconds[i].cuda(), g, hist[0], hist[1], hist[2], sch_g.coefficients(i), cfg_star=star)
        hist = [hist[0], hist[2], hist[1]]
        assert rel_l2(xg.cpu(), xc_) < 2e-5, i
        assert rel_l2(hist[1].cpu(), x0) < 2e-5 and (i == 0 or rel_l2(hist[0].cpu(), corr) < 2e-5)
    # no-CFG variant (uncond = None) takes the same path
    y = torch.randn(1, 16, 3, 8, 10, generator=gen)
    yg, h2 = y.cuda(), [torch.zeros(1, 16, 3, 8, 10, device="cuda") for _ in range(3)]
    s2 = UniPCSchedule(2, 3.0)
    co = s2.coefficients(0)
    ops.cfg_unipc_step_(yg, conds[0].cuda(), None, 1.0, h2[0], h2[1], h2[2], co)
    want, _, _ = wan_oracle.unipc_step(y.double(), conds[0].double(), 0, 0, 0, co)
    assert rel_l2(yg.cpu(), want) < 2e-6


def test_denoise_loop_unipc_matches_oracle():
    """WanDenoiser(sample_solver="unipc"): WanGP's default solver through the fused step, vs the oracle loop."""
    from oracle import wan_oracle
    from wan2gp_b200.pipeline import UniPCSchedule, WanDenoiser
    from wan2gp_b200.wan import WanModel
    cfg, thw, sd, x, t, ctx, _ = wan_case("tiny")
    m1 = WanModel(**cfg)
    m1.load_state_dict(sd)
    steps, shift, g1 = 4, 5.0, 4.0
    den = WanDenoiser(m1, num_steps=steps, shift=shift, guide_scale=g1, sample_solver="unipc")
    ctx_null = torch.zeros_like(ctx)
    got = den.generate(ctx, ctx_null, (16,) + thw, seed=11, decode=False)["latents"].cpu()
    sch = UniPCSchedule(steps, shift)
    lat = torch.randn(1, 16, *thw, generator=torch.Generator().manual_seed(11))
    xl, m0, m1_ = torch.zeros_like(lat), torch.zeros_like(lat), torch.zeros_like(lat)
    for i in range(steps):
        tt = torch.tensor([float(sch.timesteps[i])])
        c = wan_oracle.wan_forward(sd, cfg, lat, tt, ctx, emulate_bf16=True)
        u = wan_oracle.wan_forward(sd, cfg, lat, tt, ctx_null, emulate_bf16=True)
        lat, corr, x0 = wan_oracle.unipc_step(lat, wan_oracle.cfg_combine(c, u, g1), xl, m0, m1_, sch.coefficients(i))
        xl, m0, m1_ = corr, x0, m0
    r = rel_l2(got, lat)
    print(f"4-step UniPC denoise loop: latents vs oracle loop rel-L2 {r:.3e}")
    assert r < 1e-2


def test_dpmpp_through_fused_kernel():
    """sample_solver="dpm++": the same fused kernel with DPMppSchedule's coefficients, vs the fp64 restatement."""
    from oracle import wan_oracle
    from wan2gp_b200 import ops
    from wan2gp_b200.pipeline import DPMppSchedule, WanDenoiser
    steps, shift, g = 6, 5.0, 3.0
    gen = torch.Generator().manual_seed(4)
    x = torch.randn(1, 16, 2, 6, 8, generator=gen)
    sch_g, sch_c = DPMppSchedule(steps, shift), DPMppSchedule(steps, shift)
    xg, hist = x.cuda(), [torch.zeros(1, 16, 2, 6, 8, device="cuda") for _ in range(3)]
    xc, xl, m0, m1 = x.double(), 0, torch.zeros_like(x).double(), torch.zeros_like(x).double()
    for i in range(steps):
        c, u = torch.randn(x.shape, generator=gen), torch.randn(x.shape, generator=gen)
        xc, _, x0 = wan_oracle.unipc_step(xc, wan_oracle.cfg_combine(c.double(), u.double(), g), xl, m0, m1, sch_c.coefficients(i))
        m0, m1 = x0, m0
        ops.cfg_unipc_step_(xg, c.cuda(), u.cuda(), g, hist[0], hist[1], hist[2], sch_g.coefficients(i))
        hist = [hist[0], hist[2], hist[1]]
        assert rel_l2(xg.cpu(), xc) < 2e-5, i
    assert WanDenoiser(None, num_steps=4, shift=5.0, sample_solver="dpm++").timesteps[:4] == [float(t) for t in DPMppSchedule(4, 5.0).timesteps]


@pytest.mark.parametrize("solver", ["euler", "unipc", "dpm++"])
@pytest.mark.parametrize("star", [False, True])
def test_whole_step_graph_equals_eager_steps(star, solver):
    """both CFG forwards + combine + Euler update of a step as ONE captured CUDA graph, replayed with the
    timestep / guidance / dt read from device memory -- bit-identical to the launch-by-launch step over a whole schedule with an expert
    switch (two graphs) and, with star, the CFG-Zero* rescale switching on after the first step (a second capture per expert).  The
    multi-step solvers (UniPC with its corrector, dpm++) read their ten coefficients from device memory and capture once per step parity
    (the two x0 history buffers swap roles every step)."""
    from wan2gp_b200.pipeline import WanDenoiser
    from wan2gp_b200.wan import WanModel
    cfg, thw, sd, x, t, ctx, _ = wan_case("tiny")
    m1, m2 = WanModel(**cfg), WanModel(**cfg)
    m1.load_state_dict(sd), m2.load_state_dict(synth.make_wan_state_dict(cfg, seed=5))
    ctx, ctx_null = ctx.cuda(), torch.zeros_like(ctx).cuda()
    kw = dict(num_steps=6, shift=5.0, guide_scale=4.0, guide2_scale=3.0, switch_threshold=600, cfg_star_switch=star, cfg_zero_step=0,
              sample_solver=solver)
    eager, graph = WanDenoiser(m1, m2, **kw), WanDenoiser(m1, m2, **kw)
    graph.use_step_graph = True
    lat0 = torch.randn(1, 16, *thw, generator=torch.Generator().manual_seed(3)).cuda()
    a, b = lat0.clone(), lat0.clone()
    for i in range(6):
        assert eager.step(a, i, ctx, ctx_null) is a
        assert graph.step(b, i, ctx, ctx_null) is b
        assert torch.equal(a, b), i
    assert 2 <= len(graph._step_graphs) <= 8                           # one capture per (expert, CFG-Zero* phase[, step parity]) actually visited
    a2, b2 = lat0.clone(), lat0.clone()                                # a second schedule on new latents: the solver history restarts
    for i in range(3):
        eager.step(a2, i, ctx, ctx_null), graph.step(b2, i, ctx, ctx_null)
        assert torch.equal(a2, b2), i
    graph._interrupt = True                                            # the poll lives at the step boundary
    assert graph.step(b, 0, ctx, ctx_null) is None
