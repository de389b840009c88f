"""GPU edge cases of the hot path: ragged / minimal shapes, batch > 1, single-frame decode, frame cropping, error behaviour."""
import math

import pytest
import torch

from tests.helpers import rel_l2, vae_case, wan_case
from wan2gp_b200 import synth

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]
bf16 = torch.bfloat16


class Pipe:
    _interrupt = False


def test_attention_ragged_minimal():
    from wan2gp_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    for Lq, Lk, H in [(1, 1, 1), (3, 129, 2), (257, 7, 1), (255, 511, 3), (8, 8, 5)]:
        q, k, v = (torch.randn(L, H * 128, device="cuda", generator=g).to(bf16) for L in (Lq, Lk, Lk))
        out = ops.attention(q, k, v, H)
        qh, kh, vh = (t.double().reshape(-1, H, 128).permute(1, 0, 2) for t in (q, k, v))
        ref = (torch.softmax(qh @ kh.transpose(1, 2) / math.sqrt(128), -1) @ vh).permute(1, 0, 2).reshape(Lq, -1)
        assert torch.isfinite(out).all() and rel_l2(out, ref) < 4e-3


def test_wan_batch2_and_single_frame():
    """x entries with B = 2 (any2video.py:1470 batch_size) and a single latent frame (image generation)."""
    from oracle import wan_oracle
    from wan2gp_b200.wan import WanModel
    cfg, thw, sd, x, t, ctx, y = wan_case("tiny")
    m = WanModel(**cfg)
    m.load_state_dict(sd)
    x2 = torch.cat([x, x.flip(2)], 0)
    out = m([x2.clone()], t, [ctx], pipeline=Pipe())[0].cpu()
    ref = wan_oracle.wan_forward(sd, cfg, x2, t, ctx, emulate_bf16=True)
    assert out.shape == ref.shape == (2, 16) + thw and rel_l2(out, ref) < 5e-3
    x1 = x[:, :, :1, :6, :10].contiguous()                      # T=1, H=6, W=10 -> L = 15 tokens
    out1 = m([x1.clone()], t, [ctx], pipeline=Pipe())[0].cpu()  # freqs=None -> tables computed by the product code
    assert rel_l2(out1, wan_oracle.wan_forward(sd, cfg, x1, t, ctx, emulate_bf16=True)) < 5e-3


def test_vae_single_frame_odd_sizes_and_crop():
    """Tl = 1 (no temporal up-sampling at all), spatial sizes that are not multiples of the 8x16 conv tile, frame cropping."""
    from oracle import vae_oracle
    from wan2gp_b200.wan import WanVAE
    cfg, sd, _ = vae_case("vae_tiny")
    vae = WanVAE(device="cuda", state_dict=sd, cfg=cfg)
    for shape in [(16, 1, 5, 7), (16, 2, 3, 9)]:
        z = synth._normal(shape, 1.0, 7, "edge.z", "cpu")
        got = vae.model.decode(z[None].cuda(), vae.scale)[0].cpu()
        ref = vae_oracle.vae_decode(sd, z, synth.VAE_MEAN, synth.VAE_STD, cfg, emulate_bf16=True)
        assert got.shape == ref.shape == (3, 4 * (shape[1] - 1) + 1, 8 * shape[2], 8 * shape[3])
        assert rel_l2(got, ref) < 2.5e-2
    z = synth._normal((16, 3, 4, 6), 1.0, 8, "edge.z2", "cpu")
    full = vae.decode_to_cpu_uint8([z.cuda()], 0)[0]
    crop = vae.decode_to_cpu_uint8([z.cuda()], 0, target_frames=4, target_height=24, target_width=40, frame_start=2)[0]
    assert crop.shape == (3, 4, 24, 40) and torch.equal(crop, full[:, 2:6, :24, :40])


def test_error_behaviour():
    from wan2gp_b200 import _lib, ops
    from wan2gp_b200.wan import WanModel, WanVAE
    cfg, thw, sd, x, t, ctx, y = wan_case("tiny")
    m = WanModel(**cfg)
    m.load_state_dict(sd)
    with pytest.raises(NotImplementedError):
        m([x.clone()], t, [ctx], vace_context=[x], pipeline=Pipe())
    with pytest.raises(NotImplementedError):
        m([x.clone()], torch.tensor([1.0, 2.0, 3.0]), [ctx], pipeline=Pipe())
    with pytest.raises(_lib.B200Error):
        ops.gemm(torch.zeros(4, 12, device="cuda", dtype=bf16), torch.zeros(8, 12, device="cuda", dtype=bf16))     # K % 8
    with pytest.raises(RuntimeError):                                   # no weights loaded: loud failure, no fallback
        WanVAE(device="cuda").decode([torch.zeros(16, 1, 2, 2)], tile_size=256)
    with pytest.raises(NotImplementedError):
        WanVAE(device="cuda").decode([torch.zeros(16, 1, 2, 2)], tile_size=0, any_end_frame=True)


@pytest.mark.parametrize("C", [16, 96, 384, 1024])
def test_rms_silu_widths(C):
    """RMS_norm + SiLU over channels up to the 1024-channel levels of the Hunyuan 1.5 VAE, vs a plain torch fp32 reference."""
    from wan2gp_b200.wan.vae import rms_silu
    g = torch.Generator(device="cuda").manual_seed(C)
    x = (torch.randn(3, 5, 7, C, device="cuda", generator=g) * 1.5).to(bf16)
    gam = 1 + 0.1 * torch.randn(C, device="cuda", generator=g)
    for silu in (True, False):
        y = rms_silu(x, gam, silu=silu).float()
        ref = torch.nn.functional.normalize(x.float(), dim=-1) * C ** 0.5 * gam
        ref = torch.nn.functional.silu(ref) if silu else ref
        assert rel_l2(y, ref) < 4e-3


@pytest.mark.parametrize("C,G", [(32, 8), (64, 32), (128, 32), (512, 32)])
def test_group_norm_cl(C, G):
    """Clip-wide GroupNorm statistics + fused normalise/SiLU/replicate-pad slice writer vs torch.nn.functional.group_norm."""
    import torch.nn.functional as F
    from wan2gp_b200.hyvideo.vae10 import _GroupNorm
    g = torch.Generator(device="cuda").manual_seed(C + G)
    T, H, W = 5, 6, 7
    x = (torch.randn(T, H, W, C, device="cuda", generator=g) * 2 + 0.7).to(bf16)
    w, b = 1 + 0.1 * torch.randn(C, device="cuda", generator=g), 0.1 * torch.randn(C, device="cuda", generator=g)
    gn = _GroupNorm(w, b, G, "cuda")
    st = gn.stats(x)
    assert torch.equal(st, gn.stats(x))                                   # fixed summation order: bit-reproducible
    xc = x.float().permute(3, 0, 1, 2)[None]                               # [1,C,T,H,W]
    ref = F.group_norm(xc, G, w, b, 1e-6)
    mean, var = xc.reshape(G, -1).mean(1), xc.reshape(G, -1).var(1, unbiased=False)
    scale = (w.reshape(G, -1) / (var[:, None] + 1e-6).sqrt()).reshape(-1)
    assert torch.allclose(st[:C], scale, atol=1e-5, rtol=1e-4)                       # scale = gamma * rstd
    assert torch.allclose(st[C:], b - (mean[:, None] * scale.reshape(G, -1)).reshape(-1), atol=1e-4, rtol=1e-4)
    y = gn.apply(x, st, False).float().permute(3, 0, 1, 2)[None]
    assert rel_l2(y, ref) < 4e-3
    # SiLU + replicate padding of the time slice [2, 5): 2 frames in front come from frames 0..1, 1 pixel around
    yp = gn.apply(x, st, True, 2, 3, (2, 1, 1)).float().permute(3, 0, 1, 2)[None]
    refp = F.pad(F.silu(ref), (1, 1, 1, 1, 0, 0), mode="replicate")
    assert yp.shape == (1, C, 5, H + 2, W + 2) and rel_l2(yp, refp) < 4e-3
    # slice starting at frame 0: the 2 front frames replicate frame 0
    y0 = gn.apply(x, st, True, 0, 2, (2, 1, 1)).float().permute(3, 0, 1, 2)[None]
    ref0 = F.pad(F.silu(ref[:, :, :2]), (1, 1, 1, 1, 2, 0), mode="replicate")
    assert rel_l2(y0, ref0) < 4e-3


@pytest.mark.parametrize("mode,F,N,C", [(0, 2, 100, 64), (1, 3, 60, 128), (2, 3, 60, 64), (1, 2, 25700, 64)])
def test_attention_1head_modes(mode, F, N, C):
    """b200_attention_1head: per-frame (Wan VAE), frame-causal (Hunyuan VAEs) and full attention, N % 8 != 0, and key rows longer
    than the shared-memory softmax buffer (> 51200 keys -> the two-pass online softmax) -- vs torch fp32 softmax(QK^T)V."""
    from wan2gp_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(F * N + C)
    L = F * N
    buf = torch.zeros(L + 8, 3 * C, device="cuda", dtype=bf16)
    buf[:L] = torch.randn(L, 3 * C, device="cuda", generator=g).to(bf16)
    lk_max = L if mode else N
    npad = (lk_max + 63) // 64 * 64
    ws = torch.empty(N * npad * 6, device="cuda", dtype=torch.uint8)
    out = torch.empty(L, C, device="cuda", dtype=bf16)
    _lib.call("b200_attention_1head", buf.data_ptr(), out.data_ptr(), ws.data_ptr(), ws.numel(), F, N, C, C ** -0.5, mode,
              torch.cuda.current_stream().cuda_stream)
    q, k, v = (buf[:L, i * C:(i + 1) * C].float() for i in range(3))
    for f in range(F):
        lo, hi = (f * N, (f + 1) * N) if mode == 0 else (0, (f + 1) * N) if mode == 1 else (0, L)
        ref = torch.softmax(q[f * N:(f + 1) * N] @ k[lo:hi].t() * C ** -0.5, -1) @ v[lo:hi]
        assert rel_l2(out[f * N:(f + 1) * N].float(), ref) < 8e-3, (mode, f)
