"""CPU tests for hot-path row H5 (HunyuanVideo 1.0 VAE decode): oracle vs the reference fixtures, and the host-side algebra of
the phase-decomposed up-sampling conv (no GPU, no compute calls into the C ABI)."""
import pytest
import torch
import torch.nn.functional as F

from tests.helpers import load_golden, rel_l2
from wan2gp_b200 import synth


@pytest.mark.parametrize("name,zshape,seed", [("hyvae10_tiny", (8, 3, 2, 3), 0), ("hyvae10_small", (16, 2, 2, 2), 1)])
def test_hyvae10_oracle_matches_reference(name, zshape, seed):
    """fp32 oracle == reference AutoencoderKLCausal3D.decode (un-tiled) on the committed fixture."""
    from oracle import hyvae10_oracle
    cfg = synth.HYVAE10_CONFIGS[name]
    sd = synth.make_hyvae10_state_dict(cfg, seed)
    z = synth._normal((1,) + zshape, 1.0, seed, "input.z", "cpu")[0]
    g = load_golden(name)["out"][0]
    assert rel_l2(hyvae10_oracle.hyvae10_decode(sd, cfg, z), g) < 1e-5


def test_hyvae10_layout_884():
    """Up-sampling plan of the 884 decoder (vae/vae.py:259-261): x(1,2,2), x(2,2,2), x(2,2,2), none."""
    blocks, c_last = synth.hyvae10_layout(synth.HYVAE10_CONFIGS["hyvae10"])
    assert [up for _, up in blocks] == [(False, True), (True, True), (True, True), None]
    assert [rs[0] for rs, _ in blocks] == [(512, 512), (512, 512), (512, 256), (256, 128)] and c_last == 128


def _reference_upsample(x, w, b, up_t):
    """UpsampleCausal3D.forward (unet_causal_3d_blocks.py:196-222) in plain torch: x [C,T,H,W]."""
    if up_t:
        first = F.interpolate(x[None, :, :1], scale_factor=(1, 2, 2), mode="nearest")[0]
        x = torch.cat([first, F.interpolate(x[None, :, 1:], scale_factor=(2, 2, 2), mode="nearest")[0]], 1) if x.shape[1] > 1 else first
    else:
        x = F.interpolate(x[None], scale_factor=(1, 2, 2), mode="nearest")[0]
    return F.conv3d(F.pad(x[None], (1, 1, 1, 1, 2, 0), mode="replicate"), w, b)[0]


def _phase_convs_cpu(up, x):
    """What _UpConvNearest.__call__ asks b200_conv3d_cl_view to do, restated with F.conv3d (same windows, strides, weights)."""
    C, T, H, W = x.shape
    ptf = 1 if up.up_t else 2
    xp = F.pad(x[None], (1, 1, 1, 1, ptf, 0), mode="replicate")[0]
    out = torch.zeros(up.cout, 2 * T - 1 if up.up_t else T, 2 * H, 2 * W)
    for (pt, py, px), w in up.w.items():
        kt = 3 if pt is None else 2
        wk = w.float().reshape(up.cout, kt, 2, 2, up.cin).permute(0, 4, 1, 2, 3)
        n_t, off_t, o_t, s_t = (T, 0, 0, 1) if pt is None else (T, 0, 0, 2) if pt == 0 else (T - 1, 1, 1, 2)
        if n_t <= 0:
            continue
        win = xp[:, off_t:off_t + n_t + kt - 1, py:py + H + 1, px:px + W + 1]
        out[:, o_t::s_t, py::2, px::2][:, :n_t] = F.conv3d(win[None], wk, up.b)[0]
    return out


@pytest.mark.parametrize("up_t", [False, True])
@pytest.mark.parametrize("T", [1, 2, 5])
def test_upsample_phase_decomposition(up_t, T):
    """nearest x(1|2,2,2) + replicate-padded causal 3x3x3 conv == 4|8 phase convs with pre-summed taps on the low-res tensor."""
    from wan2gp_b200.hyvideo.vae10 import _UpConvNearest
    g = torch.Generator().manual_seed(T + 10 * up_t)
    w, b, x = torch.randn(16, 8, 3, 3, 3, generator=g) * 0.2, torch.randn(16, generator=g), torch.randn(8, T, 3, 4, generator=g)
    up = _UpConvNearest(w, b, up_t, "cpu", dtype=torch.float32)
    assert len(up.w) == (8 if up_t else 4)
    assert rel_l2(_phase_convs_cpu(up, x), _reference_upsample(x, w, b, up_t)) < 1e-5


def test_post_quant_fold():
    """conv_in(replicate_pad(post_quant_conv(z))) == folded conv (vae10.HYVAE10Decoder.load_state_dict)."""
    g = torch.Generator().manual_seed(3)
    zc, c0 = 8, 16
    wq, bq = torch.randn(zc, zc, 1, 1, 1, generator=g) * 0.3, torch.randn(zc, generator=g)
    wi, bi = torch.randn(c0, zc, 3, 3, 3, generator=g) * 0.1, torch.randn(c0, generator=g)
    z = torch.randn(1, zc, 3, 4, 5, generator=g)
    pad = lambda t: F.pad(t, (1, 1, 1, 1, 2, 0), mode="replicate")                                    # noqa: E731
    want = F.conv3d(pad(F.conv3d(z, wq, bq)), wi, bi)
    wf, bf = torch.einsum("omtyx,mi->oityx", wi, wq.reshape(zc, zc)), bi + torch.einsum("omtyx,m->o", wi, bq)
    assert rel_l2(F.conv3d(pad(z), wf, bf), want) < 1e-5


@pytest.mark.parametrize("name,cfg_name,xshape,seed", [("hyvae_enc_tiny", "hyvae_tiny", (3, 5, 16, 24), 2), ("hyvae_enc_small", "hyvae_small", (3, 5, 32, 48), 3)])
def test_hyvae15_encode_oracle_matches_reference(name, cfg_name, xshape, seed):
    """Hunyuan 1.5 VAE Encoder ('HY encode'): fp32 oracle == reference Encoder.forward on the committed fixture."""
    from oracle import hyvae_oracle
    cfg = synth.HYVAE_CONFIGS[cfg_name]
    sd = synth.make_hyvae_state_dict(cfg, seed, encoder=True)
    x = synth._normal((1,) + xshape, 0.5, seed, "input.video", "cpu").clamp_(-1, 1)[0]
    g = load_golden(name)["out"][0]
    out = hyvae_oracle.hyvae_encode(sd, cfg, x)
    assert out.shape == g.shape and rel_l2(out, g) < 5e-6


@pytest.mark.parametrize("name,cfg_name,xshape,seed", [("hyvae10_enc_tiny", "hyvae10_tiny", (3, 5, 16, 24), 2), ("hyvae10_enc_small", "hyvae10_small", (3, 9, 16, 16), 3)])
def test_hyvae10_encode_oracle_matches_reference(name, cfg_name, xshape, seed):
    """HunyuanVideo 1.0 VAE encode: fp32 oracle == reference AutoencoderKLCausal3D.encode (moments) on the committed fixture."""
    from oracle import hyvae10_oracle
    cfg = synth.HYVAE10_CONFIGS[cfg_name]
    sd = synth.make_hyvae10_state_dict(cfg, seed, encoder=True)
    x = synth._normal((1,) + xshape, 0.5, seed, "input.video", "cpu").clamp_(-1, 1)[0]
    g = load_golden(name)["out"][0]
    out = hyvae10_oracle.hyvae10_encode(sd, cfg, x)
    assert out.shape == g.shape and rel_l2(out, g) < 5e-6


@pytest.mark.parametrize("st_t", [False, True])
@pytest.mark.parametrize("T", [1, 5])
def test_strided_downsample_as_window_convs(st_t, T):
    """DownsampleCausal3D's stride-(1|2,2,2) replicate-padded conv == stride-1 window convs over the space-to-depth (and frame-pair)
    view of the padded tensor (_DownConvRep), restated with F.conv3d."""
    from tests.test_vae_enc_cpu import _view_conv_cpu
    from wan2gp_b200.hyvideo.vae10 import _DownConvRep
    g = torch.Generator().manual_seed(10 * T + st_t)
    H, W, C, Co = 4, 6, 8, 16
    x = torch.randn(T, H, W, C, generator=g)
    w, b = torch.randn(Co, C, 3, 3, 3, generator=g) * 0.2, torch.randn(Co, generator=g)
    dc = _DownConvRep(w, b, st_t, "cpu", dtype=torch.float32)
    xp = F.pad(x.permute(3, 0, 1, 2)[None], (1, 1, 1, 1, 2, 0), mode="replicate")
    ref = F.conv3d(xp, w, b, stride=(2 if st_t else 1, 2, 2))[0].permute(1, 2, 3, 0)
    xpc = xp[0].permute(1, 2, 3, 0)                                                     # [T+2, H+2, W+2, C]
    hs, ws = (H + 2) // 2, (W + 2) // 2
    s2d = xpc.reshape(T + 2, hs, 2, ws, 2, C).permute(0, 1, 3, 2, 4, 5).reshape(T + 2, hs, ws, 4 * C)
    if not st_t:
        got = _view_conv_cpu(s2d, (0, 0, 0), dc.w_all, dc.b, T, H // 2, W // 2, (3, 2, 2))
    else:
        if (T + 2) % 2:
            s2d = torch.cat([s2d, torch.full((1, hs, ws, 4 * C), float("nan"))], 0)     # spare frame, never read
        pairs = s2d.reshape(s2d.shape[0] // 2, 2 * hs, ws, 4 * C)
        to = (T - 1) // 2 + 1
        got = (_view_conv_cpu(pairs, (0, 0, 0), dc.w_even, dc.b, to, H // 2, W // 2, (2, 2, 2))
               + _view_conv_cpu(pairs, (0, hs, 0), dc.w_odd, None, to, H // 2, W // 2, (1, 2, 2)))
    assert got.shape == ref.shape and torch.isfinite(got).all() and rel_l2(got, ref) < 1e-5


@pytest.mark.parametrize("name", ["hyvae_tiled", "hyvae10_tiled"])
def test_hunyuan_tiled_decode_oracle_matches_reference(name):
    """Temporal + spatial tiling with cross-faded seams (what the reference pipelines always enable): restatement == the reference's
    AutoencoderKLConv3D / AutoencoderKLCausal3D .decode after enable_tiling()."""
    from oracle import hyvae10_oracle, hyvae_oracle
    g = load_golden(name)
    if name == "hyvae_tiled":
        cfg = synth.HYVAE_CONFIGS["hyvae_tiny"]
        sd, sde = synth.make_hyvae_state_dict(cfg, 6), synth.make_hyvae_state_dict(cfg, 6, encoder=True)
        z = synth._normal((1, 8, 7, 6, 10), 1.0, 6, "input.z", "cpu")[0]
        xv = synth._normal((1, 3, 13, 24, 40), 0.5, 6, "input.video", "cpu").clamp_(-1, 1)[0]
        fn = lambda t: hyvae_oracle.hyvae_decode(sd, cfg, t)                                          # noqa: E731
        fe = lambda t: hyvae_oracle.hyvae_encode(sde, cfg, t)                                         # noqa: E731
    else:
        cfg = synth.HYVAE10_CONFIGS["hyvae10_tiny"]
        sd = synth.make_hyvae10_state_dict(cfg, 7, encoder=True)
        z = synth._normal((1, 8, 7, 5, 7), 1.0, 7, "input.z", "cpu")[0]
        xv = synth._normal((1, 3, 25, 40, 56), 0.5, 7, "input.video", "cpu").clamp_(-1, 1)[0]
        fn = lambda t: hyvae10_oracle.hyvae10_decode(sd, cfg, t)                                      # noqa: E731
        fe = lambda t: hyvae10_oracle.hyvae10_encode(sd, cfg, t)                                      # noqa: E731
    tiles = (int(g["lat_size"]), int(g["lat_tsize"]), int(g["sample_size"]), int(g["sample_tsize"]))
    out = hyvae_oracle.tiled_decode(fn, z, *tiles)
    assert out.shape == g["out"][0].shape and rel_l2(out, g["out"][0]) < 5e-6
    enc = hyvae_oracle.tiled_encode(fe, xv, *tiles)                                                     # tiled encode -> posterior moments
    assert enc.shape == g["enc"][0].shape and rel_l2(enc, g["enc"][0]) < 5e-6
