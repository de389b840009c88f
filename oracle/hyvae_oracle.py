"""CPU oracle for the Hunyuan Video 1.5 VAE decoder (TEST INFRASTRUCTURE ONLY) -- hot-path row H6 of SURVEY.md section 8a:
models/hyvideo/vae/hunyuanvideo_15_vae.py  Decoder.forward (:486-520), ResnetBlock (:217-250), AttnBlock (:184-214) with the
frame-causal mask of prepare_causal_attention_mask (:161-181), Upsample (:299-339), CausalConv3d with REPLICATE padding
(:124-158), RMS_norm (:107-122).  Un-tiled decode (AutoencoderKLConv3D.decode with tiling off, :889-907).

Pinned by tests/golden/hyvae_*.npz = outputs of the reference Decoder (oracle/gen_golden.py).  The decoder config is an
external download (hunyuan.py:329-336); fixtures use reduced configs and the upstream channel plan.
"""
import torch
import torch.nn.functional as F

from wan2gp_b200.synth import hyvae_encoder_layout, hyvae_layout


def _q(t, on):
    return t.to(torch.bfloat16).to(torch.float32) if on else t


def causal_conv3d_rep(x, w, b, em=False):
    """x [C,T,H,W]; replicate padding: (k-1) frames in FRONT in time, k//2 each side in space (:137-158)."""
    k = w.shape[2]
    if k > 1:
        x = F.pad(x[None], (k // 2, k // 2, k // 2, k // 2, k - 1, 0), mode="replicate")[0]
    return F.conv3d(_q(x, em)[None], _q(w.float(), em), b.float())[0]


def rms_norm_c(x, gamma):
    n = x.norm(dim=0, keepdim=True).clamp_min(1e-12)
    return x / n * (x.shape[0] ** 0.5) * gamma.float().reshape(-1, 1, 1, 1)


def resnet(sd, p, x, em):
    h = _q(F.silu(rms_norm_c(x, sd[p + "norm1.gamma"])), em)
    h = _q(causal_conv3d_rep(h, sd[p + "conv1.conv.weight"], sd[p + "conv1.conv.bias"], em), em)
    h = _q(F.silu(rms_norm_c(h, sd[p + "norm2.gamma"])), em)
    h = causal_conv3d_rep(h, sd[p + "conv2.conv.weight"], sd[p + "conv2.conv.bias"], em)
    if p + "nin_shortcut.weight" in sd:
        x = _q(causal_conv3d_rep(_q(x, em), sd[p + "nin_shortcut.weight"], sd[p + "nin_shortcut.bias"], em), em)
    return _q(h + x, em)


def attn_block(sd, p, x, em):
    """Full 3-D single-head attention, key frame <= query frame (:161-214)."""
    C, T, H, W = x.shape
    hn = _q(rms_norm_c(x, sd[p + "norm.gamma"]), em)
    tok = hn.permute(1, 2, 3, 0).reshape(T * H * W, C)
    q, k, v = (_q(tok @ _q(sd[p + n + ".weight"].float().reshape(C, C), em).t() + sd[p + n + ".bias"].float(), em) for n in ("q", "k", "v"))
    s = (q @ k.t()) / (C ** 0.5)
    frame = torch.arange(T * H * W) // (H * W)
    s = s.masked_fill(frame[None, :] > frame[:, None], float("-inf"))
    pr = torch.exp(s - s.max(-1, keepdim=True).values)
    o = _q((_q(pr, em) @ v) / pr.sum(-1, keepdim=True), em)
    o = o @ _q(sd[p + "proj_out.weight"].float().reshape(C, C), em).t() + sd[p + "proj_out.bias"].float()
    return _q(x + o.reshape(T, H, W, C).permute(3, 0, 1, 2), em)


def upsample(sd, p, x, cout, temporal, em):
    """Upsample (:299-339): conv to factor*cout channels, channel -> (time,) space shuffle, repeat-interleave shortcut."""
    Cin, T, H, W = x.shape
    factor = 8 if temporal else 4
    rep = factor * cout // Cin
    h = _q(causal_conv3d_rep(x, sd[p + "conv.conv.weight"], sd[p + "conv.conv.bias"], em), em)

    def shuf(t, r1):            # "(r1 r2 r3 c) f h w -> c (f r1) (h r2) (w r3)"
        c = t.shape[0] // (r1 * 4)
        f = t.shape[1]
        return t.reshape(r1, 2, 2, c, f, H, W).permute(3, 4, 0, 5, 1, 6, 2).reshape(c, f * r1, 2 * H, 2 * W)
    if temporal:
        h_first = shuf(h[:, :1], 1)
        h_first = h_first[: h_first.shape[0] // 2]
        hh = torch.cat([h_first, shuf(h[:, 1:], 2)], 1) if T > 1 else h_first
        x_first = shuf(x[:, :1], 1).repeat_interleave(rep // 2, dim=0)
        sc = torch.cat([x_first, shuf(x[:, 1:], 2).repeat_interleave(rep, dim=0)], 1) if T > 1 else x_first
    else:
        hh = shuf(h, 1)
        sc = shuf(x.repeat_interleave(rep, dim=0), 1)
    return _q(hh + sc, em)


def hyvae_decode(sd, cfg, z, emulate_bf16=False):
    """z [zc,T,h,w] fp32 -> frames [3, ft*(T-1)+1, fs*h, fs*w] fp32 (Decoder.forward, :486-520)."""
    em = emulate_bf16
    c0 = cfg["block_out_channels"][0]
    z = _q(z.float(), em)
    h = causal_conv3d_rep(z, sd["conv_in.conv.weight"], sd["conv_in.conv.bias"], em) + z.repeat_interleave(c0 // cfg["z_channels"], dim=0)
    h = _q(h, em)
    h = resnet(sd, "mid.block_1.", h, em)
    h = attn_block(sd, "mid.attn_1.", h, em)
    h = resnet(sd, "mid.block_2.", h, em)
    levels, _ = hyvae_layout(cfg)
    for i, (blocks, up) in enumerate(levels):
        for j in range(len(blocks)):
            h = resnet(sd, f"up.{i}.block.{j}.", h, em)
        if up is not None:
            h = upsample(sd, f"up.{i}.upsample.", h, up[1], up[2], em)
    h = _q(F.silu(rms_norm_c(h, sd["norm_out.gamma"])), em)
    return causal_conv3d_rep(h, sd["conv_out.conv.weight"], sd["conv_out.conv.bias"], em)


def downsample(sd, p, x, cout, temporal, em):
    """Downsample (:253-296): conv to cout/factor channels, space(-time) -> channel shuffle, group-mean shortcut of the shuffled input."""
    Cin, T, H, W = x.shape
    h = _q(causal_conv3d_rep(x, sd[p + "conv.conv.weight"], sd[p + "conv.conv.bias"], em), em)

    def shuf(t, r1):            # "c (f r1) (h r2) (w r3) -> (r1 r2 r3 c) f h w"
        c, f = t.shape[0], t.shape[1] // r1
        return t.reshape(c, f, r1, H // 2, 2, W // 2, 2).permute(2, 4, 6, 0, 1, 3, 5).reshape(r1 * 4 * c, f, H // 2, W // 2)

    def gmean(t, g):
        return t.reshape(cout, g, *t.shape[1:]).mean(1)
    factor = 8 if temporal else 4
    gs = factor * Cin // cout
    if temporal:
        hf = shuf(h[:, :1], 1)
        hh = torch.cat([torch.cat([hf, hf], 0), shuf(h[:, 1:], 2)], 1) if T > 1 else torch.cat([hf, hf], 0)
        xf = gmean(shuf(x[:, :1], 1), gs // 2)
        sc = torch.cat([xf, gmean(shuf(x[:, 1:], 2), gs)], 1) if T > 1 else xf
    else:
        hh, sc = shuf(h, 1), gmean(shuf(x, 1), gs)
    return _q(hh + sc, em)


def hyvae_encode(sd, cfg, x, emulate_bf16=False):
    """x [3, 1+4k, H, W] fp32 -> moments [2 zc, 1+k, H/fs, W/fs] fp32 = (mean, logvar) of the posterior (Encoder.forward, :395-430)."""
    em = emulate_bf16
    h = _q(causal_conv3d_rep(_q(x.float(), em), sd["conv_in.conv.weight"], sd["conv_in.conv.bias"], em), em)
    levels, c_mid = hyvae_encoder_layout(cfg)
    for i, (blocks, down) in enumerate(levels):
        for j in range(len(blocks)):
            h = resnet(sd, f"down.{i}.block.{j}.", h, em)
        if down is not None:
            h = downsample(sd, f"down.{i}.downsample.", h, down[1], down[2], em)
    h = resnet(sd, "mid.block_1.", h, em)
    h = attn_block(sd, "mid.attn_1.", h, em)
    h = resnet(sd, "mid.block_2.", h, em)
    zc2 = 2 * cfg["z_channels"]
    sc = h.reshape(zc2, c_mid // zc2, *h.shape[1:]).mean(1)                    # "b (c r) f h w -> b c r f h w" mean over r
    y = _q(F.silu(rms_norm_c(h, sd["norm_out.gamma"])), em)
    return causal_conv3d_rep(y, sd["conv_out.conv.weight"], sd["conv_out.conv.bias"], em) + _q(sc, em)


def tiled(fn, x, in_s, in_t, out_s, out_t, overlap=0.25, spatial=True, temporal=True):
    """The tiling dispatch shared by both Hunyuan VAEs, decode and encode (hunyuanvideo_15_vae.py:650-703, 806-864, 866-896;
    autoencoder_kl_causal_3d.py:435-482, 597-855): temporal tiles of in_t+1 frames with stride int(in_t*(1-overlap)) (first output frame
    of every tile but the first dropped, blend_t over int(out_t*overlap) frames), each mapped through spatial tiles of in_s with stride
    int(in_s*(1-overlap)) (blend_v / blend_h over int(out_s*overlap)).  Decode: in = latent tile sizes, out = sample tile sizes; encode:
    the other way round.  fn: [C,T,h,w] -> [C',T',h',w']."""
    from .vae_oracle import spatial_tiles

    def sp(t):
        if spatial and (t.shape[-1] > in_s or t.shape[-2] > in_s):
            blend = int(out_s * overlap)
            return spatial_tiles(t, in_s, int(in_s * (1 - overlap)), fn, blend, out_s - blend)
        return fn(t)
    if not (temporal and x.shape[1] > in_t):
        return sp(x)
    stride, blend = int(in_t * (1 - overlap)), int(out_t * overlap)
    t_limit = out_t - blend
    row = []
    for i in range(0, x.shape[1], stride):
        d = sp(x[:, i:i + in_t + 1]).clone()
        row.append(d[:, 1:] if i > 0 else d)
    out = []
    for i, t in enumerate(row):
        if i > 0:
            a = row[i - 1]
            e = min(a.shape[1], t.shape[1], blend)
            for k in range(e):
                t[:, k] = a[:, -e + k] * (1 - k / e) + t[:, k] * (k / e)
            out.append(t[:, :t_limit])
        else:
            out.append(t[:, :t_limit + 1])
    return torch.cat(out, 1)


def tiled_decode(decode_fn, z, lat_size, lat_tsize, sample_size, sample_tsize, **kw):
    return tiled(decode_fn, z, lat_size, lat_tsize, sample_size, sample_tsize, **kw)


def tiled_encode(encode_fn, x, lat_size, lat_tsize, sample_size, sample_tsize, **kw):
    return tiled(encode_fn, x, sample_size, sample_tsize, lat_size, lat_tsize, **kw)
