"""GPU parity of the WanVAE decode (B200 kernels through the reference-shaped WanVAE API) against the oracle and
the golden frames produced by the UNMODIFIED reference's chunked decode.  Tolerances: vs the bf16-emulating oracle
rel-L2 <= 2.5e-2 (two decorrelated bf16 paths through ~35 layers; each is ~1.2e-2 from fp32); vs the reference fp32 frames PSNR >= 35 dB on [-1,1] data (peak 2) and mean |uint8 diff| <= 1.5."""
import pytest
import torch

from tests.helpers import load_golden, psnr, rel_l2, vae_case
from wan2gp_b200 import synth

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]
bf16 = torch.bfloat16


def test_conv3d_vs_torch():
    import torch.nn.functional as F
    from wan2gp_b200.wan.vae import _Conv
    g = torch.Generator(device="cuda").manual_seed(0)
    for (T, H, W, ci, co, k) in [(3, 9, 20, 64, 128, (3, 3, 3)), (2, 8, 16, 96, 96, (3, 3, 3)), (4, 5, 7, 32, 64, (3, 1, 1)),
                                 (1, 12, 18, 192, 96, (1, 3, 3)), (2, 6, 6, 16, 384, (3, 3, 3)), (3, 8, 8, 128, 3, (3, 3, 3)), (2, 9, 17, 96, 3, (3, 3, 3)), (3, 10, 33, 96, 96, (3, 3, 3)),
                                 (2, 4, 4, 128, 256, (1, 1, 1))]:
        x = torch.randn(T, H, W, ci, device="cuda", generator=g).to(bf16)
        w = (torch.randn(co, ci, *k, device="cuda", generator=g) * (ci * k[0] * k[1] * k[2]) ** -0.5)
        b = torch.randn(co, device="cuda", generator=g)
        conv = _Conv(w, b, "cuda")
        xp = F.pad(x.float().permute(3, 0, 1, 2)[None], (k[2] // 2, k[2] // 2, k[1] // 2, k[1] // 2, k[0] - 1, 0))
        ref = F.conv3d(xp.double(), w.to(bf16).double(), b.double())[0]          # [co, T, H, W]
        if co == 3:
            out = conv(x, out_mode=2)
            assert rel_l2(out, ref) < 1e-3
        else:
            r = torch.randn(T, H, W, co, device="cuda", generator=g).to(bf16)
            out = conv(x, residual=r)
            assert rel_l2(out.permute(3, 0, 1, 2), ref + r.double().permute(3, 0, 1, 2)) < 4e-3


def test_conv3d_row_kernel_vs_torch():
    """Image rows >= ~103 pixels take the row-tiled kernel (csrc/conv_sm100.cuh: one halo tile per frame tap, all kh x kw taps
    from shifted shared-memory descriptors).  Zero-padded (Wan) and replicate-padded (Hunyuan) convs, ragged W / odd H, two N
    tiles, Cin not a multiple of 64, the planar fp32 head, the residual epilogue -- vs torch conv3d in fp64."""
    import torch.nn.functional as F
    from wan2gp_b200.hyvideo.vae import _RepConv
    from wan2gp_b200.wan.vae import _Conv
    g = torch.Generator(device="cuda").manual_seed(5)
    for (T, H, W, ci, co, k) in [(3, 5, 130, 64, 128, (3, 3, 3)), (2, 3, 256, 96, 96, (3, 3, 3)), (2, 4, 300, 128, 3, (3, 3, 3)),
                                 (2, 5, 128, 256, 256, (1, 3, 3)), (2, 7, 200, 16, 384, (3, 3, 3)), (1, 2, 330, 32, 64, (3, 3, 3)),
                                 (4, 1, 110, 128, 32, (3, 3, 3))]:
        x = torch.randn(T, H, W, ci, device="cuda", generator=g).to(bf16)
        w = (torch.randn(co, ci, *k, device="cuda", generator=g) * (ci * k[0] * k[1] * k[2]) ** -0.5)
        b = torch.randn(co, device="cuda", generator=g)
        xc = x.float().permute(3, 0, 1, 2)[None]
        pads = (k[2] // 2, k[2] // 2, k[1] // 2, k[1] // 2, k[0] - 1, 0)
        for cls, mode in ((_Conv, "constant"), (_RepConv, "replicate")):
            conv = cls(w, b, "cuda")
            ref = F.conv3d(F.pad(xc, pads, mode=mode).double(), w.to(bf16).double(), b.double())[0]          # [co, T, H, W]
            if co == 3:
                assert rel_l2(conv(x, out_mode=2), ref) < 1e-3, (cls.__name__, T, H, W, ci, co)
            else:
                r = torch.randn(T, H, W, co, device="cuda", generator=g).to(bf16)
                out = conv(x, residual=r)
                assert rel_l2(out.permute(3, 0, 1, 2), ref + r.double().permute(3, 0, 1, 2)) < 4e-3, (cls.__name__, T, H, W, ci, co)


def test_upsample_convs_row_kernel():
    """Phase-decomposed up-sampling convs (Wan 2x2 sub-pixel; Hunyuan 1.0 nearest x(1|2,2,2) with 2x2(x2) phases) on rows wide
    enough for the row-tiled kernel."""
    import torch.nn.functional as F
    from wan2gp_b200.hyvideo.vae10 import _UpConvNearest
    from wan2gp_b200.wan.vae import _UpConv
    g = torch.Generator(device="cuda").manual_seed(6)
    T, H, W, ci, co = 3, 5, 136, 64, 32
    x = torch.randn(T, H, W, ci, device="cuda", generator=g).to(bf16)
    w2 = torch.randn(co, ci, 3, 3, device="cuda", generator=g) * (ci * 9) ** -0.5
    b = torch.randn(co, device="cuda", generator=g)
    out = _UpConv(w2, b, "cuda")(x)
    up = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=(2.0, 2.0), mode="nearest-exact")
    assert rel_l2(out, F.conv2d(up.double(), w2.double(), b.double(), padding=1).permute(0, 2, 3, 1)) < 5e-3
    w3 = torch.randn(co, ci, 3, 3, 3, device="cuda", generator=g) * (ci * 27) ** -0.5
    xc = x.float().permute(3, 0, 1, 2)[None]
    for up_t in (False, True):
        out = _UpConvNearest(w3, b, up_t, "cuda")(x)
        if up_t:
            u = torch.cat([F.interpolate(xc[:, :, :1], scale_factor=(1, 2, 2), mode="nearest"),
                           F.interpolate(xc[:, :, 1:], scale_factor=(2, 2, 2), mode="nearest")], 2)
        else:
            u = F.interpolate(xc, scale_factor=(1, 2, 2), mode="nearest")
        ref = F.conv3d(F.pad(u, (1, 1, 1, 1, 2, 0), mode="replicate").double(), w3.double(), b.double())[0]
        assert rel_l2(out.permute(3, 0, 1, 2), ref) < 6e-3, up_t


@pytest.mark.parametrize("name", ["vae_tiny", "vae_small", "vae_p"])
def test_vae_decode(name):
    from oracle import vae_oracle
    from wan2gp_b200.wan import WanVAE
    cfg, sd, z = vae_case(name)
    vae = WanVAE(device="cuda", state_dict=sd, cfg=cfg)
    got = vae.model.decode(z[None].cuda(), vae.scale)[0].cpu()
    g = load_golden(name)["out"][0]
    assert got.shape == g.shape
    line = f"{name}: vs reference frames rel-L2 {rel_l2(got, g):.3e}, PSNR {psnr(got.clamp(-1, 1), g.clamp(-1, 1), 2.0):.1f} dB"
    if name != "vae_p":
        emu = vae_oracle.vae_decode(sd, z, synth.VAE_MEAN, synth.VAE_STD, cfg, emulate_bf16=True)
        line += f"; vs bf16-emulating oracle {rel_l2(got, emu):.3e}"
        assert rel_l2(got, emu) < 2.5e-2
    u8 = vae.decode_to_cpu_uint8([z.cuda()], 0)[0]
    ref8 = vae_oracle.frames_to_uint8(g)
    d = (u8.int() - ref8.int()).abs()
    print(line + f"; uint8 mean|d| {d.float().mean():.3f} max {int(d.max())}")
    assert u8.dtype == torch.uint8 and u8.device.type == "cpu" and u8.shape == ref8.shape
    assert psnr(got.clamp(-1, 1), g.clamp(-1, 1), 2.0) > 35.0
    assert d.float().mean() <= 1.5
    fl = vae.decode([z.cuda()], 0)[0]
    assert fl.min() >= -1 and fl.max() <= 1


def test_upconv_subpixel_vs_torch():
    """nearest-exact 2x + Conv2d 3x3 (vae.py:124-133) computed as four 2x2 sub-pixel convs on the low-res input."""
    import torch.nn.functional as F
    from wan2gp_b200.wan.vae import _UpConv
    g = torch.Generator(device="cuda").manual_seed(1)
    for (T, H, W, ci, co) in [(2, 5, 7, 64, 32), (1, 9, 20, 192, 96), (3, 8, 16, 384, 192)]:
        x = torch.randn(T, H, W, ci, device="cuda", generator=g).to(bf16)
        w = torch.randn(co, ci, 3, 3, device="cuda", generator=g) * (ci * 9) ** -0.5
        b = torch.randn(co, device="cuda", generator=g)
        out = _UpConv(w, b, "cuda")(x)
        up = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=(2.0, 2.0), mode="nearest-exact")
        ref = F.conv2d(up.double(), w.double(), b.double(), padding=1).permute(0, 2, 3, 1)
        assert out.shape == (T, 2 * H, 2 * W, co)
        assert rel_l2(out, ref) < 5e-3


@pytest.mark.parametrize("name", ["vae_enc_tiny", "vae_enc_small", "vae_enc_1f"])
def test_vae_encode(name):
    """Wan VAE encode through the reference-shaped WanVAE.encode: whole-sequence causal convs, stride-2 convs as
    space-to-depth / frame-pair window convs, folded head.  vs the reference's chunked encode (fixture) and the bf16-emulating
    oracle: rel-L2 <= 2.5e-2 (the emulating oracle is 5e-3 from the fp32 reference)."""
    from oracle import vae_oracle
    from tests.test_vae_enc_cpu import enc_case
    from wan2gp_b200.wan import WanVAE
    cfg, sd, x = enc_case(name)
    vae = WanVAE(state_dict=sd, cfg=cfg)
    got = vae.encode([x[0].cuda()], tile_size=0)[0].cpu()
    g = load_golden(name)["out"][0]
    emu = vae_oracle.vae_encode(sd, x[0], synth.VAE_MEAN, synth.VAE_STD, cfg, emulate_bf16=True)
    print(f"{name}: vs reference {rel_l2(got, g):.3e}; vs bf16-emulating oracle {rel_l2(got, emu):.3e}")
    assert got.shape == g.shape and got.dtype == torch.float32
    assert rel_l2(got, g) < 2.5e-2 and rel_l2(got, emu) < 2.5e-2


def test_vae_encode_wide_rows_and_roundtrip():
    """Rows wide enough for the row-tiled conv kernel at the first two levels (W = 416 -> 208 -> 104), vs the bf16-emulating oracle;
    then decode(encode(x)) runs end to end with the expected shapes (1 + 4k frames -> 1 + k latent frames -> 1 + 4k frames)."""
    from oracle import vae_oracle
    from wan2gp_b200.wan import WanVAE
    cfg = synth.VAE_CFG_TINY
    sd = synth.make_vae_state_dict(cfg, 3, encoder=True)
    x = synth._normal((3, 5, 16, 416), 0.5, 3, "input.video", "cpu").clamp_(-1, 1)
    vae = WanVAE(state_dict=sd, cfg=cfg)
    mu = vae.encode([x.cuda()], tile_size=0)[0]
    emu = vae_oracle.vae_encode(sd, x, synth.VAE_MEAN, synth.VAE_STD, cfg, emulate_bf16=True)
    print(f"wide encode: vs bf16-emulating oracle {rel_l2(mu.cpu(), emu):.3e}")
    assert mu.shape == (16, 2, 2, 52) and rel_l2(mu.cpu(), emu) < 2.5e-2
    rec = vae.decode([mu])[0]
    assert rec.shape == (3, 5, 16, 416) and torch.isfinite(rec).all()
    with pytest.raises(ValueError):
        vae.encode([x.cuda()], tile_size=20)
    with pytest.raises(ValueError):
        vae.encode([x[:, :4].cuda()], tile_size=0)


def test_vae_tiled_decode_encode():
    """tile_size > 0: latent / video tiles with 25 % overlap, each run through the whole-clip decoder / encoder, seams
    cross-faded by b200_blend_edge_f32 -- vs the reference's spatial_tiled_decode, its streaming tiled uint8 writer and
    spatial_tiled_encode (fixtures)."""
    from wan2gp_b200.wan import WanVAE
    cfg = synth.VAE_CFG_TINY
    g = load_golden("vae_tiled_dec")
    tile = int(g["tile"])
    vae = WanVAE(state_dict=synth.make_vae_state_dict(cfg, 4, encoder=True), cfg=cfg)
    z = synth._normal((1, 16, 2, 12, 14), 1.0, 4, "input.z", "cpu")[0]
    got = vae.decode([z.cuda()], tile_size=tile)[0].cpu()
    ref = g["out"][0].clamp(-1, 1)
    print(f"tiled decode: vs reference rel-L2 {rel_l2(got, ref):.3e}, PSNR {psnr(got, ref, 2.0):.1f} dB")
    assert got.shape == ref.shape and psnr(got, ref, 2.0) > 35.0 and rel_l2(got, ref) < 2.5e-2
    u8 = vae.decode_to_cpu_uint8([z.cuda()], tile_size=tile, target_frames=4, target_height=90, target_width=100, frame_start=1)[0]
    want = torch.from_numpy(g["u8"][0])[:, 1:5, :90, :100]
    d = (u8.int() - want.int()).abs().float()
    print(f"tiled uint8: mean |d| {d.mean():.3f} max {d.max():.0f}")
    assert u8.shape == want.shape and d.mean() < 1.5
    # a single tile (tile_size >= frame) must equal the un-tiled decode bit for bit
    assert torch.equal(vae.decode([z.cuda()], tile_size=256)[0], vae.decode([z.cuda()], tile_size=0)[0])
    g = load_golden("vae_tiled_enc")
    vae = WanVAE(state_dict=synth.make_vae_state_dict(cfg, 5, encoder=True), cfg=cfg)
    x = synth._normal((1, 3, 5, 96, 112), 0.5, 5, "input.video", "cpu").clamp_(-1, 1)[0]
    mu = vae.encode([x.cuda()], tile_size=int(g["tile"]))[0].cpu()
    print(f"tiled encode: vs reference rel-L2 {rel_l2(mu, g['out'][0]):.3e}")
    assert mu.shape == g["out"][0].shape and rel_l2(mu, g["out"][0]) < 2.5e-2


def test_streamed_decode_equals_whole_clip():
    """Time-sliced decode (per-conv 2-frame history carried between slices = the reference's feature-cache decode, vae.py:639-655, for
    slices of several latent frames) against the whole-clip decode: same kernels on the same operands, so bit-identical.  Exercises
    the fused-norm epilogues, the tap-stacked head with history frames, time_conv across slice boundaries and ragged last slices."""
    from wan2gp_b200.wan import WanVAE
    sd = synth.make_vae_state_dict(synth.VAE_CFG, 3)
    vae = WanVAE(device="cuda", state_dict=sd)
    for (T, h, w), chunk in (((7, 16, 26), 3), ((10, 30, 52), 4), ((9, 16, 32), 8)):
        z = synth._normal((16, T, h, w), 1.0, T, "input.zs", "cpu").cuda()
        whole = vae.model.decode_frames(z, vae.mean, vae.std)
        sliced = vae.model.decode_frames_streamed(z, vae.mean, vae.std, chunk=chunk)
        assert whole.shape == sliced.shape == (3, 4 * (T - 1) + 1, 8 * h, 8 * w)
        assert bool(torch.isfinite(whole).all())
        d = (whole - sliced).abs().max()
        print(f"streamed vs whole decode T={T} chunk={chunk}: max|d| = {float(d):.3e}")
        assert float(d) == 0.0
