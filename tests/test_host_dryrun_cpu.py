"""CPU: dry-run of the host-side VAE modules with the C ABI stubbed out -- no kernel runs (and nothing is computed: outputs are
uninitialised), but every `_lib.call` must pass exactly as many arguments as the ctypes signature declares and all tensor shapes the
host code derives (slices, strides, spare frames, tiles) must be consistent.  The arithmetic is covered by the `-m gpu` parity tests."""
import pytest
import torch

from wan2gp_b200 import _lib, ops, synth


@pytest.fixture
def stub_abi(monkeypatch):
    import wan2gp_b200.hyvideo.vae as hv
    import wan2gp_b200.hyvideo.vae10 as hv10
    import wan2gp_b200.wan.vae as wv
    calls = []

    def fake_call(name, *args):
        assert len(args) == len(_lib.SIGNATURES[name]), (name, len(args), len(_lib.SIGNATURES[name]))
        calls.append(name)
        return 0
    monkeypatch.setattr(_lib, "call", fake_call)
    for m in (wv, hv, hv10):
        monkeypatch.setattr(m, "_s", lambda: 0)
    monkeypatch.setattr(ops, "_stream", lambda: 0)
    monkeypatch.setattr(ops, "_chk", lambda *a, **k: None)
    return calls


def test_hunyuan_vae_host_paths(stub_abi):
    import wan2gp_b200.hyvideo.vae as hv
    import wan2gp_b200.hyvideo.vae10 as hv10
    cfg = synth.HYVAE10_CONFIGS["hyvae10_tiny"]
    sd = synth.make_hyvae10_state_dict(cfg, 0, encoder=True)
    enc, dec = hv10.HYVAE10Encoder(cfg, "cpu"), hv10.HYVAE10Decoder(cfg, "cpu")
    enc.load_state_dict(sd), dec.load_state_dict(sd)
    assert enc(torch.randn(1, 3, 5, 16, 24)).shape == (1, 16, 2, 2, 3)
    assert dec(torch.randn(1, 8, 2, 2, 3)).shape == (1, 3, 5, 16, 24)
    assert {"b200_group_stats_cl", "b200_group_norm_apply_cl", "b200_space_to_depth_cl", "b200_conv3d_cl_view"} <= set(stub_abi)
    cfg = synth.HYVAE_CONFIGS["hyvae_tiny"]
    enc, dec = hv.HYVAEEncoder(cfg, "cpu"), hv.HYVAEDecoder(cfg, "cpu")
    enc.load_state_dict(synth.make_hyvae_state_dict(cfg, 0, encoder=True)), dec.load_state_dict(synth.make_hyvae_state_dict(cfg, 0))
    assert enc(torch.randn(1, 3, 5, 16, 24)).shape == (1, 16, 3, 4, 6)
    assert dec(torch.randn(1, 8, 3, 4, 6)).shape == (1, 3, 5, 16, 24)
    assert {"b200_hy_downsample_cl", "b200_group_mean_cl", "b200_hy_upsample_cl", "b200_rms_silu_pad_cl"} <= set(stub_abi)
    # tiled dispatch (enable_tiling): 7 latent frames -> temporal tiles of 4(+1) with stride 3, spatial tiles of 4 with stride 3
    from wan2gp_b200.hyvideo import AutoencoderKLCausal3D, AutoencoderKLConv3D
    vae = AutoencoderKLConv3D(latent_channels=8, block_out_channels=[32, 64, 64], layers_per_block=1, ffactor_spatial=4, ffactor_temporal=2,
                              sample_size=16, sample_tsize=8, device="cpu")
    full = {"decoder." + k: v for k, v in synth.make_hyvae_state_dict(cfg, 0).items()}
    full.update({"encoder." + k: v for k, v in synth.make_hyvae_state_dict(cfg, 0, encoder=True).items()})
    vae.load_state_dict(full)
    vae.enable_tiling()
    assert vae.decode(torch.randn(1, 8, 7, 6, 10), return_dict=False)[0].shape == (1, 3, 13, 24, 40)
    assert vae.encode(torch.randn(1, 3, 13, 24, 40)).latent_dist.mean.shape == (1, 8, 7, 6, 10)
    cfg10 = synth.HYVAE10_CONFIGS["hyvae10_tiny"]
    vae = AutoencoderKLCausal3D(sample_size=32, sample_tsize=16, device="cpu", **cfg10)
    vae.load_state_dict(synth.make_hyvae10_state_dict(cfg10, 0, encoder=True))
    vae.enable_tiling()
    assert vae.decode(torch.randn(1, 8, 7, 5, 7), return_dict=False)[0].shape == (1, 3, 25, 40, 56)
    assert vae.encode(torch.randn(1, 3, 25, 40, 56)).latent_dist.mean.shape == (1, 8, 7, 5, 7)
    assert "b200_blend_edge_f32" in stub_abi
    vae.disable_tiling()
    assert vae.decode(torch.randn(1, 8, 2, 2, 3), return_dict=True).sample.shape == (1, 3, 5, 16, 24)


def test_wan_vae_host_paths(stub_abi):
    import wan2gp_b200.wan.vae as wv
    cfg = synth.VAE_CFG_TINY
    vae = wv.WanVAE(device="cpu", state_dict=synth.make_vae_state_dict(cfg, 0, encoder=True), cfg=cfg)
    assert vae.encode([torch.randn(3, 5, 32, 48)], tile_size=0)[0].shape == (16, 2, 4, 6)
    assert vae.encode([torch.randn(3, 5, 96, 112)], tile_size=64)[0].shape == (16, 2, 12, 14)
    assert vae.decode([torch.randn(16, 2, 12, 14)], tile_size=64)[0].shape == (3, 5, 96, 112)
    assert vae.decode([torch.randn(16, 2, 4, 6)], tile_size=0)[0].shape == (3, 5, 32, 48)
    assert {"b200_blend_edge_f32", "b200_planar_to_cl_pad", "b200_upconv2x_cl", "b200_vae_prologue"} <= set(stub_abi)


class _Pipe:
    _interrupt = False


def test_dit_host_paths(stub_abi):
    """WanModel (t2v, i2v joint CFG pair with the prompt cache) and HYVideoDiffusionTransformer (1.5 and 1.0 families): every launch of a
    whole forward has a well-formed argument list and the host-side shape algebra closes."""
    from wan2gp_b200.hyvideo import HYVideoDiffusionTransformer
    from wan2gp_b200.wan import WanModel
    thw = (3, 8, 12)
    cfg = synth.WAN_CONFIGS["tiny"]
    m = WanModel(**cfg, device="cpu")
    m.load_state_dict(synth.make_wan_state_dict(cfg, 0))
    x, t, ctx, y = synth.make_wan_inputs(cfg, thw, 0)
    assert m([x.clone()], t, [ctx], pipeline=_Pipe())[0].shape == (1, 16) + thw
    cfg = synth.WAN_CONFIGS["tiny_i2v"]
    m = WanModel(**cfg, device="cpu")
    m.load_state_dict(synth.make_wan_state_dict(cfg, 0))
    x, t, ctx, y = synth.make_wan_inputs(cfg, thw, 0)
    m.cache_context = True
    ctx0 = ctx * 0
    for _ in range(2):                                        # second call is served from the prompt cache
        outs = m([x.clone(), x.clone()], t, [ctx, ctx0], y=y, pipeline=_Pipe())
        assert [o.shape for o in outs] == [(1, 16) + thw] * 2
    assert len(m._prompts.emb) == 2 and len(m._prompts.ckv) == 2 * len(m.blocks)
    n_kv = stub_abi.count("b200_rmsnorm_rope")
    m([x.clone(), x.clone()], t, [ctx, ctx0], y=y, pipeline=_Pipe())
    per_forward_cached = stub_abi.count("b200_rmsnorm_rope") - n_kv
    m.cache_context = False
    n_kv = stub_abi.count("b200_rmsnorm_rope")
    m([x.clone(), x.clone()], t, [ctx, ctx0], y=y, pipeline=_Pipe())
    assert stub_abi.count("b200_rmsnorm_rope") - n_kv == per_forward_cached + 2 * len(m.blocks)      # the cross-K norms come back
    for name in ("hy_tiny", "hy10_tiny"):
        cfg = synth.HY_CONFIGS[name]
        v10 = cfg.get("family") == "1.0"
        kw = dict(mm_single_blocks_depth=cfg["mm_single_blocks_depth"], text_states_dim_2=cfg["text_states_dim_2"], guidance_embed=True) if v10 else dict(mm_single_blocks_depth=0, text_pool_type=None, glyph_byT5_v2=True, use_cond_type_embedding=True, pre_split_qkv=True)
        hm = HYVideoDiffusionTransformer(i2v_condition_type=None, patch_size=cfg["patch_size"], in_channels=cfg["in_channels"],
                                         out_channels=cfg["out_channels"], hidden_size=cfg["hidden_size"], heads_num=cfg["heads_num"],
                                         mm_double_blocks_depth=cfg["mm_double_blocks_depth"], text_states_dim=cfg["text_states_dim"],
                                         device="cpu", **kw)
        hm.load_state_dict(synth.make_hy_state_dict(cfg, 0))
        thw2 = (2, 8, 12) if v10 else (3, 6, 10)
        xi, tt, txt, tm, b5, bm = synth.make_hy_inputs(cfg, thw2, seed=0)
        extra = dict(text_states_2=torch.randn(1, cfg["text_states_dim_2"]), guidance=torch.tensor([6000.0])) if v10 else dict(byt5_text_states=b5, byt5_text_mask=bm)
        out = hm(xi, tt, text_states=txt, text_mask=tm, pipeline=_Pipe(), **extra)
        assert out.shape == (1, cfg["out_channels"]) + thw2


@pytest.mark.parametrize("solver", ["euler", "unipc", "dpm++", "lcm", "causvid"])
def test_denoiser_host_paths(stub_abi, solver):
    """WanDenoiser.step for every sample solver: schedule bookkeeping, history rotation and the ctypes coefficient block of the fused step."""
    from wan2gp_b200.pipeline import WanDenoiser
    from wan2gp_b200.wan import WanModel
    cfg = synth.WAN_CONFIGS["tiny"]
    m = WanModel(**cfg, device="cpu")
    m.load_state_dict(synth.make_wan_state_dict(cfg, 0))
    den = WanDenoiser(m, num_steps=3, shift=5.0, guide_scale=4.0, device="cpu", sample_solver=solver, cfg_star_switch=True, cfg_zero_step=0)
    x, t, ctx, _ = synth.make_wan_inputs(cfg, (3, 8, 12), 0)
    lat = x.clone()
    for i in range(den.num_steps):
        assert den.step(lat, i, ctx, ctx * 0) is lat
    kernel = "b200_cfg_euler_step" if solver in ("euler", "lcm", "causvid") else "b200_cfg_unipc_step"
    assert stub_abi.count(kernel) == den.num_steps
