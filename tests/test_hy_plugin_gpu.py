"""GPU: `HunyuanVideoSampler.generate(**kwargs)` obtained through the plugin's Hunyuan `family_handler.load_model` (boundary levels 1-2)
computes what the reference's generate computes for the same arguments -- text states -> CUDA-generator noise in the latent dtype ->
FlowMatchDiscreteScheduler (reverse, euler) with the CFG pair / CFG-Zero* (Hunyuan 1.5) or the embedded guidance (HunyuanVideo 1.0) -> VAE
decode -- checked against the oracle loop assembled from the same pieces (reduced architectures)."""
import pytest
import torch

from tests.helpers import psnr, rel_l2
from tests.test_hy_plugin_cpu import hy_kwargs, make_pipeline
from wan2gp_b200 import synth

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


def _states(enc, prompt, is_uncond=False):
    out = enc.encode(enc.text2tokens([prompt]), is_uncond=is_uncond)
    return out.hidden_state.to(torch.bfloat16).float(), out.attention_mask


def test_generate_hunyuan_1_5_matches_oracle_loop(monkeypatch):
    from oracle import hy_oracle, hyvae_oracle, wan_oracle
    from wan2gp_b200.pipeline import flow_match_timesteps
    pipe_obj, pipe, cfg, sd, vsd = make_pipeline("b200_hunyuan_1_5_t2v", device="cuda", monkeypatch=monkeypatch, vae_tiling=False)
    steps, shift, g, seed = 3, 9.0, 6.0, 11
    kw = hy_kwargs(sampling_steps=steps, shift=shift, guide_scale=g, seed=seed, frame_num=5, height=32, width=48, cfg_star_switch=1)
    out = pipe_obj.generate(**kw)
    assert out.dtype == torch.float32 and out.device.type == "cpu" and tuple(out.shape) == (3, 5, 32, 48)
    # ---- oracle loop: same noise call (one CUDA generator, bf16 draw: hunyuan.py:815, pipeline :1464-1467, prepare_latents :860-866)
    thw = (3, 8, 12)
    lat = torch.randn((1, 8) + thw, generator=torch.Generator("cuda").manual_seed(seed), device="cuda", dtype=torch.bfloat16).float().cpu()
    te = pipe_obj.text_encoder
    txt, tm = _states(te, kw["input_prompt"])
    txtn, tmn = _states(te, pipe_obj.default_negative_prompt, True)
    b5, bm = torch.zeros(1, 256, 1472), torch.zeros(1, 256, dtype=torch.long)            # no quoted glyph text in the prompt
    ts = flow_match_timesteps(steps, shift)
    cond_lat = torch.zeros(1, 9, *thw)
    freqs = hy_oracle.rope_tables_hy(thw)
    for i in range(steps):
        x = torch.cat([lat, cond_lat], 1)
        t = torch.tensor([ts[i]])
        c = hy_oracle.hy_forward(sd, cfg, x, t, txt, tm, b5, bm, freqs=freqs, emulate_bf16=True)
        u = hy_oracle.hy_forward(sd, cfg, x, t, txtn, tmn, b5, bm, freqs=freqs, emulate_bf16=True)
        v = wan_oracle.cfg_combine(c, u, g, cfg_star=True, step_no=i, cfg_zero_step=-1)
        lat = wan_oracle.euler_step(lat, v, ts[i] / 1000.0, ts[i + 1] / 1000.0)
    vc = pipe_obj.vae.config
    z = lat / vc.scaling_factor + vc.shift_factor
    dsd = {k[len("decoder."):]: v for k, v in vsd.items()}
    ref = hyvae_oracle.hyvae_decode(dsd, synth.HYVAE_CONFIGS["hyvae_tiny"], z[0], emulate_bf16=True)
    r, p = rel_l2(out, ref), psnr(out.clamp(-1, 1), ref.clamp(-1, 1), 2.0)
    print(f"HunyuanVideoSampler.generate (1.5, CFG pair + CFG-Zero*, 3 steps + decode) vs oracle loop: rel-L2 {r:.3e}, PSNR {p:.1f} dB")
    assert r < 4e-2 and p > 35.0
    # joint pass = the same arithmetic in one forward of batch 2; same seed -> same clip; other seed -> other clip
    joint = pipe_obj.generate(**dict(kw, joint_pass=True))
    assert rel_l2(joint, out) < 1e-5
    assert torch.equal(pipe_obj.generate(**kw), out)
    assert not torch.equal(pipe_obj.generate(**dict(kw, seed=seed + 1)), out)
    # the reference's tiled decode (enable_tiling before every decode): same clip up to the seam blending
    pipe_obj.vae_tiling = True
    tiled = pipe_obj.generate(**kw)
    assert tuple(tiled.shape) == (3, 5, 32, 48) and torch.isfinite(tiled).all()


def test_generate_hunyuan_1_0_matches_oracle_loop(monkeypatch):
    from oracle import hy_oracle, hyvae10_oracle, wan_oracle
    from wan2gp_b200.pipeline import flow_match_timesteps
    pipe_obj, pipe, cfg, sd, vsd = make_pipeline("b200_hunyuan", device="cuda", monkeypatch=monkeypatch, vae_tiling=False)
    steps, shift, seed = 3, 7.0, 5
    kw = hy_kwargs(model_type="b200_hunyuan", sampling_steps=steps, shift=shift, guide_scale=7.0, embedded_guidance_scale=6.0, seed=seed,
                   frame_num=9, height=32, width=48)
    out = pipe_obj.generate(**kw)
    assert tuple(out.shape) == (3, 9, 32, 48)
    thw = (3, 4, 6)
    lat = torch.randn((1, 8) + thw, generator=torch.Generator("cuda").manual_seed(seed), device="cuda", dtype=torch.bfloat16).float().cpu()
    txt, tm = _states(pipe_obj.text_encoder, kw["input_prompt"])
    e2 = pipe_obj.text_encoder_2
    pooled = e2.encode(e2.text2tokens([kw["input_prompt"]])).hidden_state.to(torch.bfloat16).float()
    gd = (torch.tensor([6.0]).to(torch.bfloat16) * 1000.0).float()                   # 6016: the reference multiplies in the bf16 latent dtype (:1661-1670)
    assert float(gd) == 6016.0
    ts = flow_match_timesteps(steps, shift)
    freqs = hy_oracle.rope_tables_hy((3, 2, 3))
    for i in range(steps):
        v = hy_oracle.hy_forward(sd, cfg, lat, torch.tensor([ts[i]]), txt, tm, freqs=freqs, emulate_bf16=True, text_states_2=pooled, guidance=gd)
        lat = wan_oracle.euler_step(lat, v, ts[i] / 1000.0, ts[i + 1] / 1000.0)
    z = lat / pipe_obj.vae.config.scaling_factor
    ref = hyvae10_oracle.hyvae10_decode(vsd, synth.HYVAE10_CONFIGS["hyvae10_tiny"], z[0], emulate_bf16=True)
    r, p = rel_l2(out, ref), psnr(out.clamp(-1, 1), ref.clamp(-1, 1), 2.0)
    print(f"HunyuanVideoSampler.generate (1.0, embedded guidance, 3 steps + decode) vs oracle loop: rel-L2 {r:.3e}, PSNR {p:.1f} dB")
    assert r < 8e-2 and p > 30.0
