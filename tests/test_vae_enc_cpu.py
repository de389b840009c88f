"""CPU tests for the Wan VAE encode row: oracle vs the reference fixtures, and the host-side algebra that maps the
encoder's stride-2 convs and its head onto the stride-1 conv kernels (no GPU, no compute calls into the C ABI)."""
import pytest
import torch
import torch.nn.functional as F

from tests.helpers import load_golden, rel_l2
from wan2gp_b200 import synth

ENC_CASES = {"vae_enc_tiny": (synth.VAE_CFG_TINY, (3, 9, 32, 48), 0), "vae_enc_small": (synth.VAE_CFG, (3, 5, 48, 64), 1),
             "vae_enc_1f": (synth.VAE_CFG_TINY, (3, 1, 16, 24), 2)}


def enc_case(name):
    cfg, xshape, seed = ENC_CASES[name]
    sd = synth.make_vae_state_dict(cfg, seed, encoder=True)
    x = synth._normal((1,) + xshape, 0.5, seed, "input.video", "cpu").clamp_(-1, 1)
    return cfg, sd, x


@pytest.mark.parametrize("name", list(ENC_CASES))
def test_vae_encode_oracle_matches_reference(name):
    """whole-sequence oracle == the reference's chunked (1,4,4,...) WanVAE_.encode with feature caches."""
    from oracle import vae_oracle
    cfg, sd, x = enc_case(name)
    out = vae_oracle.vae_encode(sd, x[0], synth.VAE_MEAN, synth.VAE_STD, cfg)
    g = load_golden(name)["out"][0]
    assert out.shape == g.shape and rel_l2(out, g) < 5e-6


def test_encoder_param_names_cover_reference_layout():
    s = synth.vae_encoder_param_shapes(synth.VAE_CFG)
    assert s["encoder.conv1.weight"] == (96, 3, 3, 3, 3) and s["encoder.head.2.weight"] == (32, 384, 3, 3, 3)
    kinds = [u[0] for u in synth.vae_encoder_layout(synth.VAE_CFG)[1]]
    assert kinds == ["res", "res", "down2d", "res", "res", "down3d", "res", "res", "down3d", "res", "res"]
    assert "encoder.downsamples.5.time_conv.weight" in s and "encoder.downsamples.2.time_conv.weight" not in s


def _view_conv_cpu(x, off, w_taps, bias, T, H, W, k):
    """What b200_conv3d_cl_view computes, restated with F.conv3d: x [Ti,Hi,Wi,Ci] channels-last, w_taps [Co, taps, Ci], zeros beyond
    the high end of x; returns [T,H,W,Co]."""
    kt, kh, kw = k
    Ti, Hi, Wi, Ci = x.shape
    xp = F.pad(x.permute(3, 0, 1, 2)[None], (0, kw, 0, kh, 0, kt))                    # zero fill past the end
    win = xp[:, :, off[0]:off[0] + T + kt - 1, off[1]:off[1] + H + kh - 1, off[2]:off[2] + W + kw - 1]
    wk = w_taps.float().reshape(-1, kt, kh, kw, Ci).permute(0, 4, 1, 2, 3)
    return F.conv3d(win, wk, bias)[0].permute(1, 2, 3, 0)


def test_downsample_as_space_to_depth_conv():
    """ZeroPad2d((0,1,0,1)) + Conv2d(3, stride 2) == 2x2 stride-1 conv over the space-to-depth tensor (_DownConv)."""
    from wan2gp_b200.wan.vae import _DownConv
    g = torch.Generator().manual_seed(0)
    T, H, W, C, Co = 2, 6, 8, 8, 16
    x = torch.randn(T, H, W, C, generator=g)
    w, b = torch.randn(Co, C, 3, 3, generator=g) * 0.2, torch.randn(Co, generator=g)
    dc = _DownConv(w, b, "cpu", dtype=torch.float32)
    s2d = x.reshape(T, H // 2, 2, W // 2, 2, C).permute(0, 1, 3, 2, 4, 5).reshape(T, H // 2, W // 2, 4 * C)   # channel (2p+q)*C + c
    got = _view_conv_cpu(s2d, (0, 0, 0), dc.w, dc.b, T, H // 2, W // 2, (1, 2, 2))
    ref = F.conv2d(F.pad(x.permute(0, 3, 1, 2), (0, 1, 0, 1)), w, b, stride=2).permute(0, 2, 3, 1)
    assert rel_l2(got, ref) < 1e-5


@pytest.mark.parametrize("T", [1, 3, 5, 9])
def test_time_downsample_as_frame_pair_convs(T):
    """out[0] = y[0]; out[j] = time_conv(y[2j-2], y[2j-1], y[2j]) == the reference's chunked stride-2 time_conv with its one-frame
    cache, and == two window convs over the frame-pair view (_TimeDownConv)."""
    from wan2gp_b200.wan.vae import _TimeDownConv
    g = torch.Generator().manual_seed(T)
    h, w_, C = 3, 4, 8
    y = torch.randn(T, h, w_, C, generator=g)
    w, b = torch.randn(C, C, 3, 1, 1, generator=g) * 0.3, torch.randn(C, generator=g)
    # reference chunking (vae.py:190-212): chunk 0 = 1 frame (cached, bypass), later chunks of 2 frames at this level
    yc = y.permute(3, 0, 1, 2)[None]
    outs, cache = [yc[:, :, :1]], yc[:, :, :1]
    for s in range(1, T, 2):
        chunk = yc[:, :, s:s + 2]
        outs.append(F.conv3d(torch.cat([cache[:, :, -1:], chunk], 2), w, b, stride=(2, 1, 1)))
        cache = chunk
    ref = torch.cat(outs, 2)[0].permute(1, 2, 3, 0)
    tc = _TimeDownConv(w, b, "cpu", dtype=torch.float32)
    m = (T - 1) // 2
    got = torch.empty(1 + m, h, w_, C)
    got[0] = y[0]
    if m:
        buf = torch.cat([y, torch.full((1, h, w_, C), float("nan"))], 0).reshape(m + 1, 2 * h, w_, C)   # spare frame is never read
        even = _view_conv_cpu(buf, (0, 0, 0), tc.w_even, tc.b, m, h, w_, (2, 1, 1))
        odd = _view_conv_cpu(buf, (0, h, 0), tc.w_odd, None, m, h, w_, (1, 1, 1))
        got[1:] = even + odd
    assert torch.isfinite(got).all() and rel_l2(got, ref) < 1e-5


def test_head_fold():
    """head conv -> conv1 (1x1x1) -> chunk(2)[0] -> (mu - mean) * inv_std == one conv with folded weights (WanVAEEncoder._head_conv)."""
    g = torch.Generator().manual_seed(9)
    z, c = 4, 8
    wh, bh = torch.randn(2 * z, c, 3, 3, 3, generator=g) * 0.1, torch.randn(2 * z, generator=g)
    w1, b1 = torch.randn(2 * z, 2 * z, generator=g) * 0.3, torch.randn(2 * z, generator=g)
    mean, inv_std = torch.randn(z, generator=g), torch.rand(z, generator=g) + 0.5
    x = torch.randn(1, c, 3, 5, 6, generator=g)
    pad = lambda t: F.pad(t, (1, 1, 1, 1, 2, 0))                                                       # noqa: E731
    mu = F.conv3d(F.conv3d(pad(x), wh, bh), w1.reshape(2 * z, 2 * z, 1, 1, 1), b1)[:, :z]
    want = (mu - mean.view(1, z, 1, 1, 1)) * inv_std.view(1, z, 1, 1, 1)
    m = inv_std[:, None] * w1[:z]
    wf, bf = torch.einsum("om,mctyx->octyx", m, wh), inv_std * (w1[:z] @ bh + b1[:z] - mean)
    assert rel_l2(F.conv3d(pad(x), wf, bf), want) < 1e-5


def test_tiled_oracle_matches_reference():
    """Tiled decode / encode: the tiling + seam cross-fade restatement == WanVAE_.spatial_tiled_decode /
    spatial_tiled_encode, and uint8 of the blended fp32 frames == the reference's streaming tiled writer decode_to_cpu_uint8."""
    from oracle import vae_oracle
    cfg = synth.VAE_CFG_TINY
    g = load_golden("vae_tiled_dec")
    sd = synth.make_vae_state_dict(cfg, 4, encoder=True)
    z = synth._normal((1, 16, 2, 12, 14), 1.0, 4, "input.z", "cpu")[0]
    out = vae_oracle.vae_decode_tiled(sd, z, synth.VAE_MEAN, synth.VAE_STD, int(g["tile"]), cfg)
    assert rel_l2(out, g["out"][0]) < 5e-6
    d = (vae_oracle.frames_to_uint8(out).int() - torch.from_numpy(g["u8"][0]).int()).abs()
    assert d.max() <= 1 and (d > 0).float().mean() < 1e-3
    g = load_golden("vae_tiled_enc")
    sd = synth.make_vae_state_dict(cfg, 5, encoder=True)
    x = synth._normal((1, 3, 5, 96, 112), 0.5, 5, "input.video", "cpu").clamp_(-1, 1)[0]
    assert rel_l2(vae_oracle.vae_encode_tiled(sd, x, synth.VAE_MEAN, synth.VAE_STD, int(g["tile"]), cfg), g["out"][0]) < 5e-6
