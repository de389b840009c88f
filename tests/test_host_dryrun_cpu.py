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
