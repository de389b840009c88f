"""GPU: `WanAny2V.generate(**kwargs)` obtained through the plugin's `family_handler.load_model` (boundary levels 1-2) computes what the
reference's generate computes for the same arguments -- text encode -> noise -> UniPC schedule with the two-phase guidance / expert
switch and CFG-Zero* -> VAE decode -- checked against the oracle loop assembled from the same pieces (reduced architectures)."""
import json
import os

import pytest
import torch

from tests.helpers import rel_l2
from tests.test_plugin_cpu import ROOT, fake_t5, load_handler, wgp_kwargs
from wan2gp_b200 import synth

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


def _load(arch, cfg_name, monkeypatch):
    h = load_handler()
    monkeypatch.setitem(h.ARCHS, arch, (cfg_name, True))
    cfg = synth.WAN_CONFIGS[cfg_name]
    sds = [synth.make_wan_state_dict(cfg, s) for s in (0, 5)]
    vsd = synth.make_vae_state_dict(synth.VAE_CFG_TINY, 0, encoder=True)
    model_def = json.load(open(os.path.join(ROOT, "plugin", "defaults", arch + ".json")))["model"]
    pipe_obj, pipe = h.family_handler.load_model(["hi", "lo"], arch, arch, model_def, text_encoder=fake_t5, state_dicts=sds, vae_state_dict=vsd, vae_cfg=synth.VAE_CFG_TINY)
    return pipe_obj, cfg, sds, vsd


def test_generate_t2v_matches_oracle_loop(monkeypatch):
    from oracle import vae_oracle, wan_oracle
    from wan2gp_b200.pipeline import UniPCSchedule
    pipe_obj, cfg, sds, vsd = _load("b200_t2v_2_2", "tiny", monkeypatch)
    steps, shift, g1, g2, thr, seed = 4, 5.0, 4.0, 3.0, 900, 11     # UniPC timesteps 999, 937, 833, 624: two steps per expert
    kw = wgp_kwargs(sampling_steps=steps, shift=shift, guide_scale=g1, guide2_scale=g2, switch_threshold=thr, seed=seed, frame_num=9, height=64,
                    width=96, cfg_star_switch=1, cfg_zero_step=-1)
    out = pipe_obj.generate(**kw)
    thw = (3, 8, 12)
    assert out["x"].dtype == torch.uint8 and tuple(out["x"].shape) == (3, 9, 64, 96)
    # oracle loop with the same noise (CUDA generator, any2video.py:548-549, 1470), contexts and schedule
    lat = torch.randn(1, 16, *thw, dtype=torch.float32, device="cuda", generator=torch.Generator(device="cuda").manual_seed(seed)).cpu()

    def enc(p):
        c = fake_t5([p], "cpu")[0]
        return torch.cat([c, c.new_zeros(cfg["text_len"] - c.shape[0], c.shape[1])])[None]
    ctx, ctxn = enc(kw["input_prompt"]), enc(pipe_obj.sample_neg_prompt)
    sch = UniPCSchedule(steps, shift)
    xl, m0, m1 = torch.zeros_like(lat), torch.zeros_like(lat), torch.zeros_like(lat)
    used = set()
    for i in range(steps):
        t = float(sch.timesteps[i])
        w, gs = (sds[1], g2) if t <= thr else (sds[0], g1)
        used.add(gs)
        tt = torch.tensor([t])
        c = wan_oracle.wan_forward(w, cfg, lat, tt, ctx, emulate_bf16=True)
        u = wan_oracle.wan_forward(w, cfg, lat, tt, ctxn, emulate_bf16=True)
        v = wan_oracle.cfg_combine(c, u, gs, cfg_star=True, step_no=i, cfg_zero_step=-1)
        lat, corr, x0 = wan_oracle.unipc_step(lat, v, xl, m0, m1, sch.coefficients(i))
        xl, m0, m1 = corr, x0, m0
    assert used == {g1, g2}
    ref8 = vae_oracle.frames_to_uint8(vae_oracle.vae_decode(vsd, lat[0], synth.VAE_MEAN, synth.VAE_STD, synth.VAE_CFG_TINY, emulate_bf16=True))
    d = (out["x"].int() - ref8.int()).abs().float()
    print(f"WanAny2V.generate (plugin, t2v, unipc, 2 phases, CFG-Zero*): mean |d uint8| vs oracle loop {d.mean():.3f}, max {int(d.max())}")
    assert d.mean() < 2.0
    # same seed -> same frames (noise, schedule and kernels are deterministic); a different seed -> different frames
    again = pipe_obj.generate(**kw)["x"]
    assert torch.equal(again, out["x"])
    assert not torch.equal(pipe_obj.generate(**dict(kw, seed=seed + 1))["x"], out["x"])


def test_generate_i2v_condition_matches_oracle(monkeypatch):
    from oracle import vae_oracle
    pipe_obj, cfg, sds, vsd = _load("b200_i2v_2_2", "tiny_i2v", monkeypatch)
    img = (torch.rand(3, 64, 96, generator=torch.Generator().manual_seed(3)) * 2 - 1)
    y, h, w = pipe_obj._i2v_condition(img.cuda(), None, 9, 64, 96, 0, 1.0)
    enc = torch.cat([img[:, None], torch.zeros(3, 8, 64, 96)], 1)
    ref = vae_oracle.vae_encode(vsd, enc, synth.VAE_MEAN, synth.VAE_STD, synth.VAE_CFG_TINY, emulate_bf16=True)
    r = rel_l2(y[4:].cpu(), ref)
    print(f"i2v conditioning latents vs oracle encode: rel-L2 {r:.3e}")
    assert tuple(y.shape) == (20, 3, 8, 12) and r < 2.5e-2
    out = pipe_obj.generate(**wgp_kwargs(model_type="b200_i2v_2_2", image_start=img, frame_num=9, height=64, width=96, shift=5.0,
                                         guide_scale=3.5, guide2_scale=3.5, switch_threshold=900, sampling_steps=3))
    assert tuple(out["x"].shape) == (3, 9, 64, 96) and out["x"].float().std() > 0
