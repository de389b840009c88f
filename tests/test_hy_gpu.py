"""GPU parity of the Hunyuan Video 1.5 DiT forward (B200 kernels through the reference-shaped HYVideoDiffusionTransformer
API) against the bf16-emulating oracle and the golden output of the UNMODIFIED reference (fp32 weights + the reference's
own bf16 hard-casts, see oracle/hy_oracle.py).  Tolerances: rel-L2 <= 6e-3 vs either."""
import pytest
import torch

from tests.helpers import load_golden, psnr, rel_l2
from wan2gp_b200 import synth

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


class Pipe:
    _interrupt = False


def _case():
    cfg, thw = synth.HY_CONFIGS["hy_tiny"], (3, 6, 10)
    return cfg, thw, synth.make_hy_state_dict(cfg, 0), synth.make_hy_inputs(cfg, thw, seed=0)


def _model(cfg, sd):
    from wan2gp_b200.hyvideo import HYVideoDiffusionTransformer
    m = HYVideoDiffusionTransformer(i2v_condition_type=None, patch_size=cfg["patch_size"], in_channels=cfg["in_channels"],
                                    out_channels=cfg["out_channels"], hidden_size=cfg["hidden_size"], heads_num=cfg["heads_num"],
                                    mlp_width_ratio=cfg["mlp_width_ratio"], mm_double_blocks_depth=cfg["mm_double_blocks_depth"],
                                    mm_single_blocks_depth=0, text_states_dim=cfg["text_states_dim"], text_pool_type=None,
                                    glyph_byT5_v2=True, use_cond_type_embedding=True, pre_split_qkv=True)
    m.load_state_dict(sd)
    return m


def test_hy_ops():
    """per-head RMSNorm + RoPE, SiLU / erf-GELU epilogues, pre-rounded LN modulation, patch-1 embed / unpatchify."""
    from oracle import hy_oracle, wan_oracle
    from wan2gp_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(0)
    L, H = 3 * 6 * 10, 2
    D = H * 128
    cos, sin = hy_oracle.rope_tables_hy((3, 6, 10))
    x = torch.randn(L, 3 * D, device="cuda", generator=g).to(torch.bfloat16)
    w = 1 + 0.1 * torch.randn(128, device="cuda", generator=g)
    ref = hy_oracle.rms_head(x[:, D:2 * D].float().cpu().reshape(L, H, 128), w.cpu())
    ref = wan_oracle.apply_rope(ref, cos, sin).reshape(L, D)
    ops.rmsnorm_rope_(x[:, D:2 * D], w, 1e-6, cos.cuda(), sin.cuda(), per_head=True)
    assert rel_l2(x[:, D:2 * D].cpu(), ref) < 4e-3
    a = torch.randn(70, 64, device="cuda", generator=g).to(torch.bfloat16)
    b = (torch.randn(96, 64, device="cuda", generator=g) / 8).to(torch.bfloat16)
    lin = a.double() @ b.double().t()
    assert rel_l2(ops.gemm(a, b, act=2, out_dtype=torch.float32), torch.nn.functional.silu(lin)) < 1e-3
    assert rel_l2(ops.gemm(a, b, act=3, out_dtype=torch.float32), torch.nn.functional.gelu(lin)) < 1e-3
    xf = torch.randn(50, 256, device="cuda", generator=g)
    sh, sc = torch.randn(256, device="cuda", generator=g), torch.randn(256, device="cuda", generator=g) * 0.3
    ln = torch.nn.functional.layer_norm(xf, (256,), eps=1e-6).to(torch.bfloat16).double()
    assert rel_l2(ops.ln_modulate(xf, sh, sc, pre_round=True), ln * (1 + sc.double()) + sh.double()) < 4e-3
    xin = torch.randn(65, 3, 6, 10, device="cuda", generator=g)
    wpe, bpe = torch.randn(256, 65, device="cuda", generator=g) / 8, torch.randn(256, device="cuda", generator=g)
    assert rel_l2(ops.patch_embed(xin, None, wpe, bpe, 256, patch=1), xin.reshape(65, -1).t() @ wpe.t() + bpe) < 1e-5
    y = torch.randn(L, 32, device="cuda", generator=g)
    assert torch.equal(ops.unpatchify(y, 32, 3, 6, 10, patch=1, c_major=True), y.t().reshape(32, 3, 6, 10))
    assert rel_l2(ops.col_mean(xf), xf.mean(0)) < 1e-6


def test_hy_forward_tiny():
    from oracle import hy_oracle
    cfg, thw, sd, (x, t, txt, tm, b5, bm) = _case()
    m = _model(cfg, sd)
    cos, sin = hy_oracle.rope_tables_hy(thw)
    out = m(x, t, text_states=txt, text_mask=tm, freqs_cos=cos, freqs_sin=sin, pipeline=Pipe(), byt5_text_states=b5, byt5_text_mask=bm)
    assert out.shape == (1, cfg["out_channels"]) + thw and out.dtype == torch.float32
    got = out.cpu()
    emu = hy_oracle.hy_forward(sd, cfg, x, t, txt, tm, b5, bm, emulate_bf16=True)
    g = load_golden("hy_tiny")["out"]
    print(f"hy_tiny: vs bf16-emulating oracle {rel_l2(got, emu):.3e}; vs reference {rel_l2(got, g):.3e}, "
          f"PSNR {psnr(got, g, float(g.abs().max())):.1f} dB")
    assert rel_l2(got, emu) < 6e-3 and rel_l2(got, g) < 6e-3
    # rope tables of the product code == reference tables
    from wan2gp_b200.hyvideo import get_rotary_pos_embed
    c2, s2 = get_rotary_pos_embed(thw)
    gg = load_golden("hy_tiny")
    assert torch.equal(c2[:64], gg["cos"]) and torch.equal(s2[:64], gg["sin"])
    # interrupt contract (models.py:1146-1149)

    class Stop:
        _interrupt = True
    assert m(x, t, text_states=txt, text_mask=tm, freqs_cos=cos, freqs_sin=sin, pipeline=Stop(), byt5_text_states=b5, byt5_text_mask=bm) is None


def test_hy10_forward_tiny():
    """HunyuanVideo 1.0 family: fused qkv double block + single-stream blocks (H3), pooled-text vector, guidance embed, patch 2."""
    from oracle import hy_oracle
    from wan2gp_b200.hyvideo import HYVideoDiffusionTransformer
    cfg, thw, seed = synth.HY_CONFIGS["hy10_tiny"], (2, 8, 12), 1
    sd = synth.make_hy_state_dict(cfg, seed)
    x, t, txt, tm, _, _ = synth.make_hy_inputs(cfg, thw, seed=seed)
    t2 = synth._normal((1, cfg["text_states_dim_2"]), 1.0, seed, "hy.txt2", "cpu")
    gd = torch.tensor([6000.0])
    m = HYVideoDiffusionTransformer(i2v_condition_type=None, patch_size=cfg["patch_size"], in_channels=16, out_channels=16,
                                    hidden_size=cfg["hidden_size"], heads_num=cfg["heads_num"], mm_double_blocks_depth=1,
                                    mm_single_blocks_depth=2, text_states_dim=cfg["text_states_dim"],
                                    text_states_dim_2=cfg["text_states_dim_2"], guidance_embed=True)
    m.load_state_dict(sd)
    out = m(x, t, text_states=txt, text_mask=tm, text_states_2=t2, guidance=gd, pipeline=Pipe()).cpu()
    emu = hy_oracle.hy_forward(sd, cfg, x, t, txt, tm, emulate_bf16=True, text_states_2=t2, guidance=gd)
    g = load_golden("hy10_tiny")["out"]
    print(f"hy10_tiny: vs bf16-emulating oracle {rel_l2(out, emu):.3e}; vs reference {rel_l2(out, g):.3e}")
    assert out.shape == g.shape and rel_l2(out, emu) < 6e-3 and rel_l2(out, g) < 6e-3


@pytest.mark.parametrize("name,zshape,seed", [("hyvae_tiny", (8, 3, 4, 6), 0), ("hyvae_small", (16, 3, 2, 3), 1)])
def test_hyvae_decode(name, zshape, seed):
    """Hunyuan 1.5 VAE decoder (row H6, un-tiled): replicate-padded convs, frame-causal mid attention, shuffle up-sampling.
    Tolerances as for the Wan VAE: vs the bf16-emulating oracle rel-L2 <= 2.5e-2; vs reference frames PSNR >= 35 dB."""
    from oracle import hyvae_oracle
    from wan2gp_b200.hyvideo import HYVAEDecoder
    cfg = synth.HYVAE_CONFIGS[name]
    sd = synth.make_hyvae_state_dict(cfg, seed)
    z = synth._normal((1,) + zshape, 1.0, seed, "input.z", "cpu")
    dec = HYVAEDecoder(cfg)
    dec.load_state_dict(sd)
    got = dec(z.cuda())[0].cpu()
    g = load_golden(name)["out"][0]
    emu = hyvae_oracle.hyvae_decode(sd, cfg, z[0], emulate_bf16=True)
    print(f"{name}: vs reference {rel_l2(got, g):.3e}, PSNR {psnr(got.clamp(-1, 1), g.clamp(-1, 1), 2.0):.1f} dB; vs bf16-emulating oracle {rel_l2(got, emu):.3e}")
    assert got.shape == g.shape
    assert rel_l2(got, emu) < 2.5e-2 and psnr(got.clamp(-1, 1), g.clamp(-1, 1), 2.0) > 35.0


@pytest.mark.parametrize("name,zshape,seed", [("hyvae10_tiny", (8, 3, 2, 3), 0), ("hyvae10_small", (16, 2, 2, 2), 1)])
def test_hyvae10_decode(name, zshape, seed):
    """HunyuanVideo 1.0 VAE decode (row H5, un-tiled): post_quant_conv folded into conv_in, clip-wide GroupNorm + SiLU written
    replicate-padded, phase-decomposed nearest up-sampling convs, frame-causal mid attention.  GroupNorm re-normalises 20+
    times through random weights, so bf16 rounding noise is larger than for the RMS-norm VAEs: vs the bf16-emulating oracle
    and vs the fp32 reference rel-L2 <= 6e-2 (the emulating oracle itself sits 1.1e-2 / 3.2e-2 from the reference), and frames
    PSNR >= 35 dB vs the reference."""
    from oracle import hyvae10_oracle
    from wan2gp_b200.hyvideo import HYVAE10Decoder
    cfg = synth.HYVAE10_CONFIGS[name]
    sd = synth.make_hyvae10_state_dict(cfg, seed)
    z = synth._normal((1,) + zshape, 1.0, seed, "input.z", "cpu")
    dec = HYVAE10Decoder(cfg)
    dec.load_state_dict(sd)
    got = dec(z.cuda())[0].cpu()
    g = load_golden(name)["out"][0]
    emu = hyvae10_oracle.hyvae10_decode(sd, cfg, z[0], emulate_bf16=True)
    print(f"{name}: vs reference {rel_l2(got, g):.3e}, PSNR {psnr(got.clamp(-1, 1), g.clamp(-1, 1), 2.0):.1f} dB; vs bf16-emulating oracle {rel_l2(got, emu):.3e}")
    assert got.shape == g.shape
    assert rel_l2(got, emu) < 6e-2 and rel_l2(got, g) < 6e-2 and psnr(got.clamp(-1, 1), g.clamp(-1, 1), 2.0) > 35.0


def test_hyvae10_time_slicing_and_surface():
    """The time-sliced GroupNorm->pad->conv path (PAD_SLICE_BYTES) must equal the single-slice path bit for bit, and the
    AutoencoderKLCausal3D surface (decode(z, return_dict=False)[0], enable_tiling no-op, full-attention mode) must work."""
    from wan2gp_b200.hyvideo import AutoencoderKLCausal3D, vae10
    from wan2gp_b200.hyvideo import vae as hv
    cfg = synth.HYVAE10_CONFIGS["hyvae10_tiny"]
    sd = synth.make_hyvae10_state_dict(cfg, 0)
    z = synth._normal((1, 8, 4, 3, 4), 1.0, 5, "input.z", "cpu").cuda()
    vae = AutoencoderKLCausal3D(**{k: v for k, v in cfg.items()})
    vae.load_state_dict(sd)
    vae.enable_tiling()
    whole = vae.decode(z, return_dict=False)[0]
    assert whole.shape == (1, 3, 13, 24, 32)
    old = hv.PAD_SLICE_BYTES
    try:
        hv.PAD_SLICE_BYTES = 5 * (24 + 2) * (32 + 2) * 32 * 2               # ~3 output frames per slice at full resolution
        sliced = vae.decode(z, return_dict=True).sample
        # the Hunyuan 1.5 decoder shares the sliced norm -> SiLU -> pad -> conv path
        from wan2gp_b200.hyvideo import HYVAEDecoder
        cfg15 = synth.HYVAE_CONFIGS["hyvae_tiny"]
        d15 = HYVAEDecoder(cfg15)
        d15.load_state_dict(synth.make_hyvae_state_dict(cfg15, 0))
        z15 = synth._normal((1, 8, 5, 4, 6), 1.0, 7, "input.z", "cpu").cuda()
        sliced15 = d15(z15)
    finally:
        hv.PAD_SLICE_BYTES = old
    assert torch.equal(whole, sliced)
    assert torch.equal(d15(z15), sliced15)
    # mid_block_causal_attn off: every token attends to every token -- compare with the oracle
    from oracle import hyvae10_oracle
    cfg2 = dict(cfg, mid_block_causal_attn=False)
    dec = vae10.HYVAE10Decoder(cfg2)
    dec.load_state_dict(sd)
    got = dec(z)[0].cpu()
    emu = hyvae10_oracle.hyvae10_decode(sd, cfg2, z[0].cpu(), emulate_bf16=True)
    print(f"hyvae10 full-attention: vs bf16-emulating oracle {rel_l2(got, emu):.3e}")
    assert rel_l2(got, emu) < 6e-2


@pytest.mark.parametrize("name,cfg_name,xshape,seed", [("hyvae_enc_tiny", "hyvae_tiny", (3, 5, 16, 24), 2), ("hyvae_enc_small", "hyvae_small", (3, 5, 32, 48), 3)])
def test_hyvae15_encode(name, cfg_name, xshape, seed):
    """Hunyuan 1.5 VAE encode through the AutoencoderKLConv3D surface (`.encode(x).latent_dist.mode()`): vs the reference Encoder
    (fixture) and the bf16-emulating oracle, rel-L2 <= 2.5e-2."""
    from oracle import hyvae_oracle
    from wan2gp_b200.hyvideo import AutoencoderKLConv3D
    cfg = synth.HYVAE_CONFIGS[cfg_name]
    sd = synth.make_hyvae_state_dict(cfg, seed, encoder=True)
    full = {"encoder." + k: v for k, v in sd.items()}
    full.update({"decoder." + k: v for k, v in synth.make_hyvae_state_dict(cfg, seed).items()})
    vae = AutoencoderKLConv3D(latent_channels=cfg["z_channels"], block_out_channels=list(reversed(cfg["block_out_channels"])),
                              layers_per_block=cfg["num_res_blocks"], ffactor_spatial=cfg["ffactor_spatial"], ffactor_temporal=cfg["ffactor_temporal"])
    vae.load_state_dict(full)
    x = synth._normal((1,) + xshape, 0.5, seed, "input.video", "cpu").clamp_(-1, 1)
    post = vae.encode(x.cuda()).latent_dist
    got = torch.cat([post.mean, post.logvar], 1)[0].cpu()
    g = load_golden(name)["out"][0]
    emu = hyvae_oracle.hyvae_encode(sd, cfg, x[0], emulate_bf16=True)
    print(f"{name}: vs reference {rel_l2(got, g):.3e}; vs bf16-emulating oracle {rel_l2(got, emu):.3e}")
    assert got.shape == g.shape and rel_l2(got, g) < 2.5e-2 and rel_l2(got, emu) < 2.5e-2
    assert torch.equal(post.mode(), post.mean) and post.sample(torch.Generator().manual_seed(0)).shape == post.mean.shape
    rec = vae.decode(post.mode(), return_dict=False)[0]                    # encode -> decode round trip runs end to end
    assert rec.shape == (1,) + xshape and torch.isfinite(rec).all()


@pytest.mark.parametrize("name,cfg_name,xshape,seed", [("hyvae10_enc_tiny", "hyvae10_tiny", (3, 5, 16, 24), 2), ("hyvae10_enc_small", "hyvae10_small", (3, 9, 16, 16), 3)])
def test_hyvae10_encode(name, cfg_name, xshape, seed):
    """HunyuanVideo 1.0 VAE encode through the AutoencoderKLCausal3D surface: stride-2 replicate-padded convs as window convs over the
    space-to-depth / frame-pair views, quant_conv folded into conv_out.  vs the reference moments (fixture) and the bf16-emulating
    oracle: rel-L2 <= 6e-2 (GroupNorm net, same bound as the decoder; the emulating oracle is 1.4e-2 from fp32)."""
    from oracle import hyvae10_oracle
    from wan2gp_b200.hyvideo import AutoencoderKLCausal3D
    cfg = synth.HYVAE10_CONFIGS[cfg_name]
    sd = synth.make_hyvae10_state_dict(cfg, seed, encoder=True)
    vae = AutoencoderKLCausal3D(**cfg)
    vae.load_state_dict(sd)
    x = synth._normal((1,) + xshape, 0.5, seed, "input.video", "cpu").clamp_(-1, 1)
    post = vae.encode(x.cuda()).latent_dist
    got = torch.cat([post.mean, post.logvar], 1)[0].cpu()
    g = load_golden(name)["out"][0]
    emu = hyvae10_oracle.hyvae10_encode(sd, cfg, x[0], emulate_bf16=True)
    print(f"{name}: vs reference {rel_l2(got, g):.3e}; vs bf16-emulating oracle {rel_l2(got, emu):.3e}")
    assert got.shape == g.shape and rel_l2(got, g) < 6e-2 and rel_l2(got, emu) < 6e-2
    rec = vae.decode(post.mode(), return_dict=False)[0]                    # encode -> decode round trip runs end to end
    assert rec.shape == (1,) + xshape and torch.isfinite(rec).all()


@pytest.mark.parametrize("name", ["hyvae_tiled", "hyvae10_tiled"])
def test_hunyuan_tiled_decode(name):
    """enable_tiling() on the Hunyuan VAE wrappers: temporal + spatial tiles through the whole-clip decoder, seams cross-faded on the
    GPU -- vs the reference's tiled decode (fixture), PSNR >= 35 dB and rel-L2 <= 6e-2; disable_tiling() returns to the un-tiled path."""
    from wan2gp_b200.hyvideo import AutoencoderKLCausal3D, AutoencoderKLConv3D
    g = load_golden(name)
    ss, st = int(g["sample_size"]), int(g["sample_tsize"])
    if name == "hyvae_tiled":
        cfg = synth.HYVAE_CONFIGS["hyvae_tiny"]
        vae = AutoencoderKLConv3D(latent_channels=cfg["z_channels"], block_out_channels=list(reversed(cfg["block_out_channels"])),
                                  layers_per_block=cfg["num_res_blocks"], ffactor_spatial=cfg["ffactor_spatial"],
                                  ffactor_temporal=cfg["ffactor_temporal"], sample_size=ss, sample_tsize=st)
        full = {"decoder." + k: v for k, v in synth.make_hyvae_state_dict(cfg, 6).items()}
        full.update({"encoder." + k: v for k, v in synth.make_hyvae_state_dict(cfg, 6, encoder=True).items()})
        vae.load_state_dict(full)
        z = synth._normal((1, 8, 7, 6, 10), 1.0, 6, "input.z", "cpu")
        xv = synth._normal((1, 3, 13, 24, 40), 0.5, 6, "input.video", "cpu").clamp_(-1, 1)
    else:
        cfg = synth.HYVAE10_CONFIGS["hyvae10_tiny"]
        vae = AutoencoderKLCausal3D(sample_size=ss, sample_tsize=st, **cfg)
        vae.load_state_dict(synth.make_hyvae10_state_dict(cfg, 7, encoder=True))
        z = synth._normal((1, 8, 7, 5, 7), 1.0, 7, "input.z", "cpu")
        xv = synth._normal((1, 3, 25, 40, 56), 0.5, 7, "input.video", "cpu").clamp_(-1, 1)
    assert (vae.tile_latent_min_size, vae.tile_latent_min_tsize) == (int(g["lat_size"]), int(g["lat_tsize"]))
    untiled = vae.decode(z.cuda(), return_dict=False)[0]
    vae.enable_tiling()
    got = vae.decode(z.cuda(), return_dict=False)[0][0].cpu()
    ref = g["out"][0]
    print(f"{name}: vs reference tiled decode rel-L2 {rel_l2(got, ref):.3e}, PSNR {psnr(got.clamp(-1, 1), ref.clamp(-1, 1), 2.0):.1f} dB")
    assert got.shape == ref.shape and rel_l2(got, ref) < 6e-2 and psnr(got.clamp(-1, 1), ref.clamp(-1, 1), 2.0) > 35.0
    post = vae.encode(xv.cuda()).latent_dist                                            # tiled encode (same switches)
    mom = torch.cat([post.mean, post.logvar], 1)[0].cpu()
    print(f"{name}: tiled encode vs reference rel-L2 {rel_l2(mom, g['enc'][0]):.3e}")
    assert mom.shape == g["enc"][0].shape and rel_l2(mom, g["enc"][0]) < 6e-2
    vae.disable_tiling()
    assert torch.equal(vae.decode(z.cuda(), return_dict=False)[0], untiled)
