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


def test_glyph_prompt_runs_the_byt5_encoder(monkeypatch):
    """A prompt that quotes text takes the glyph branch of _process_single_byt5_prompt (pipeline_hunyuan_video.py:1009-1041): format ->
    tokenise (padding to byt5_max_length) -> ByT5Encoder on the B200 kernels -> states + mask into every forward."""
    import types

    from oracle import t5_oracle
    from wan2gp_b200.hyvideo.byt5 import ByT5Encoder
    pipe_obj, pipe, cfg, sd, vsd = make_pipeline("b200_hunyuan_1_5_t2v", device="cuda", monkeypatch=monkeypatch, vae_tiling=False)
    tcfg = dict(synth.T5_CONFIGS["byt5_2layer"], vocab_size=400)                                  # byT5-small widths: the mapper expects 1472
    tsd = synth.make_t5_state_dict(tcfg, 4)
    pipe_obj.byt5_model = ByT5Encoder.from_state_dict(synth.t5_to_hf_t5stack_names(tsd, tcfg["num_layers"]), device="cuda")
    pipe_obj.byt5_max_length = 32
    seen = {}

    def tokenizer(text, padding=None, max_length=None, truncation=None, add_special_tokens=None, return_tensors=None):
        b = [3 + c % 380 for c in text.encode()][:max_length - 1] + [1]
        ids = torch.zeros(1, max_length, dtype=torch.long)
        ids[0, :len(b)] = torch.tensor(b)
        mask = (torch.arange(max_length) < len(b)).long()[None]
        seen["text"], seen["ids"], seen["mask"] = text, ids, mask
        return types.SimpleNamespace(input_ids=ids, attention_mask=mask)
    pipe_obj.byt5_tokenizer = tokenizer
    pipe_obj.prompt_format = types.SimpleNamespace(format_prompt=lambda texts, styles: "".join(f'Text "{t}". ' for t in texts))
    emb, mask = pipe_obj._byt5_one('a sign that says "OPEN" and “24h”')
    assert seen["text"] == 'Text "OPEN". Text "24h". ' and tuple(emb.shape) == (1, 32, 1472) and int(mask.sum()) == int(seen["mask"].sum())
    nv = int(mask.sum())
    emu = t5_oracle.t5_encode(tsd, tcfg, seen["ids"][0], seen["mask"][0], emulate_bf16=True)
    assert rel_l2(emb[0, :nv].cpu(), emu[:nv]) < 6e-3
    e0, m0 = pipe_obj._byt5_one("no glyph text here")
    assert float(e0.abs().max()) == 0.0 and int(m0.sum()) == 0 and tuple(e0.shape) == (1, 32, 1472)
    out = pipe_obj.generate(**hy_kwargs(input_prompt='a sign that says "OPEN"', sampling_steps=2, seed=2))
    plain = pipe_obj.generate(**hy_kwargs(input_prompt="a sign that says OPEN", sampling_steps=2, seed=2))
    assert tuple(out.shape) == (3, 5, 32, 48) and torch.isfinite(out).all() and not torch.equal(out, plain)


def test_hy_i2v_forward_and_generate(monkeypatch):
    """hunyuan_1_5_i2v on the GPU: (a) the transformer forward with image-encoder tokens against the reference fixture (hy_tiny_i2v) and the
    bf16-emulating oracle; (b) generate() with a start image against the oracle loop -- single-frame VAE encode (latent_dist.mode() *
    scaling_factor) into frame 0 of the concat condition, vision tokens in both CFG branches."""
    from oracle import hy_oracle, hyvae_oracle, wan_oracle
    from tests.helpers import load_golden
    from tests.test_hy_plugin_cpu import HY15_I2V_CFG
    from wan2gp_b200.hyvideo import HYVideoDiffusionTransformer
    from wan2gp_b200.pipeline import flow_match_timesteps
    # (a)
    cfg = synth.HY_CONFIGS["hy_tiny_i2v"]
    sd = synth.make_hy_state_dict(cfg, 2)
    x, t, txt, tm, b5, bm = synth.make_hy_inputs(cfg, (3, 6, 10), seed=2)
    vs = synth.make_hy_vision_states(cfg, seed=2)
    m = HYVideoDiffusionTransformer(i2v_condition_type="latent_concat", patch_size=cfg["patch_size"], in_channels=cfg["in_channels"],
                                    out_channels=cfg["out_channels"], hidden_size=cfg["hidden_size"], heads_num=cfg["heads_num"],
                                    mlp_width_ratio=cfg["mlp_width_ratio"], mm_double_blocks_depth=cfg["mm_double_blocks_depth"],
                                    mm_single_blocks_depth=0, text_states_dim=cfg["text_states_dim"], text_pool_type=None, glyph_byT5_v2=True,
                                    use_cond_type_embedding=True, pre_split_qkv=True, vision_projection="linear", vision_states_dim=cfg["vision_states_dim"])
    m.load_state_dict(sd)
    cos, sin = hy_oracle.rope_tables_hy((3, 6, 10))

    class P:
        _interrupt = False
    got = m(x, t, text_states=txt, text_mask=tm, freqs_cos=cos, freqs_sin=sin, pipeline=P(), byt5_text_states=b5, byt5_text_mask=bm, vision_states=vs).cpu()
    emu = hy_oracle.hy_forward(sd, cfg, x, t, txt, tm, b5, bm, emulate_bf16=True, vision_states=vs)
    g = load_golden("hy_tiny_i2v")["out"]
    novis = m(x, t, text_states=txt, text_mask=tm, freqs_cos=cos, freqs_sin=sin, pipeline=P(), byt5_text_states=b5, byt5_text_mask=bm).cpu()
    print(f"hy_tiny_i2v: vs bf16-emulating oracle {rel_l2(got, emu):.3e}; vs reference {rel_l2(got, g):.3e}; without the vision tokens {rel_l2(novis, g):.3e}")
    assert rel_l2(got, emu) < 6e-3 and rel_l2(got, g) < 6e-3 and rel_l2(novis, g) > 1e-2
    vis_tok = m._vision(vs[0].cuda()).cpu()
    assert rel_l2(vis_tok, hy_oracle.vision_projection(sd, vs[0], True)) < 4e-3
    # (b)
    pipe_obj, pipe, cfg2, sd2, vsd = make_pipeline("b200_hunyuan_1_5_i2v", device="cuda", monkeypatch=monkeypatch, vae_tiling=False)
    assert cfg2 == HY15_I2V_CFG
    steps, shift, gsc, seed = 2, 7.0, 6.0, 9
    img = torch.rand(3, 32, 48, generator=torch.Generator().manual_seed(5)) * 2 - 1
    kw = hy_kwargs(model_type="b200_hunyuan_1_5_i2v", image_start=img, sampling_steps=steps, shift=shift, guide_scale=gsc, seed=seed, frame_num=5)
    out = pipe_obj.generate(**kw)
    assert tuple(out.shape) == (3, 5, 32, 48)
    thw = (3, 8, 12)
    vcfg = synth.HYVAE_CONFIGS["hyvae_tiny"]
    esd = {k[len("encoder."):]: v for k, v in vsd.items() if k.startswith("encoder.")}
    dsd = {k[len("decoder."):]: v for k, v in vsd.items() if k.startswith("decoder.")}
    vc = pipe_obj.vae.config
    mom = hyvae_oracle.hyvae_encode(esd, vcfg, img[:, None], emulate_bf16=True)               # [2 z, 1, h, w]
    img_lat = mom[:8] * vc.scaling_factor
    enc_gpu = pipe_obj.vae.encode(img[None, :, None].cuda()).latent_dist.mode()[0].cpu()
    print(f"single-frame Hunyuan 1.5 VAE encode vs bf16-emulating oracle: {rel_l2(enc_gpu, mom[:8]):.3e}")
    assert rel_l2(enc_gpu, mom[:8]) < 2.5e-2
    cond_lat = torch.zeros(1, 9, *thw)
    cond_lat[0, :8, 0], cond_lat[0, 8, 0] = img_lat[:, 0], 1.0
    lat = torch.randn((1, 8) + thw, generator=torch.Generator("cuda").manual_seed(seed), device="cuda", dtype=torch.bfloat16).float().cpu()
    te, ve = pipe_obj.text_encoder, pipe_obj.vision_encoder
    txt, tm = _states(te, kw["input_prompt"])
    txtn, tmn = _states(te, pipe_obj.default_negative_prompt, True)
    vis = ve.encode_images(ve.seen[0]).last_hidden_state.to(torch.bfloat16).float()
    b5, bm = torch.zeros(1, 256, 1472), torch.zeros(1, 256, dtype=torch.long)
    ts = flow_match_timesteps(steps, shift)
    freqs = hy_oracle.rope_tables_hy(thw)
    for i in range(steps):
        xin = torch.cat([lat, cond_lat], 1)
        tt = torch.tensor([ts[i]])
        c = hy_oracle.hy_forward(sd2, cfg2, xin, tt, txt, tm, b5, bm, freqs=freqs, emulate_bf16=True, vision_states=vis)
        u = hy_oracle.hy_forward(sd2, cfg2, xin, tt, txtn, tmn, b5, bm, freqs=freqs, emulate_bf16=True, vision_states=vis)
        lat = wan_oracle.euler_step(lat, wan_oracle.cfg_combine(c, u, gsc), ts[i] / 1000.0, ts[i + 1] / 1000.0)
    ref = hyvae_oracle.hyvae_decode(dsd, vcfg, (lat / vc.scaling_factor + vc.shift_factor)[0], emulate_bf16=True)
    r, p = rel_l2(out, ref), psnr(out.clamp(-1, 1), ref.clamp(-1, 1), 2.0)
    print(f"HunyuanVideoSampler.generate (1.5 i2v, start image + vision tokens, 2 steps + decode) vs oracle loop: rel-L2 {r:.3e}, PSNR {p:.1f} dB")
    assert r < 4e-2 and p > 35.0
    other = pipe_obj.generate(**dict(kw, image_start=-img))
    assert not torch.equal(other, out)
