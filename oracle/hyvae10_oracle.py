"""CPU oracle for the HunyuanVideo 1.0 VAE decode (TEST INFRASTRUCTURE ONLY) -- hot-path row H5 of SURVEY.md section 8a:
models/hyvideo/vae/autoencoder_kl_causal_3d.py  AutoencoderKLCausal3D._decode (:474-493, un-tiled branch: post_quant_conv ->
decoder), models/hyvideo/vae/vae.py  DecoderCausal3D (:186-365), models/hyvideo/vae/unet_causal_3d_blocks.py  CausalConv3d
(:35-70, replicate padding), UpsampleCausal3D (:96-224, nearest; first frame spatial-only), ResnetBlockCausal3D (:300-493,
GroupNorm eps 1e-6 -> SiLU -> conv, 1x1 conv_shortcut), UNetMidBlockCausal3D (:606-741) with the frame-causal mask of
prepare_causal_attention_mask (:21-30).

The mid-block attention is diffusers' `Attention` (third-party, diffusers==0.36.0, not in the reference tree): single head,
GroupNorm over the token axis, q/k/v/out Linears with bias, residual add -- restated from its published behaviour; the
golden fixtures run the reference classes with the same restatement (oracle/refshim.py::load_reference_hyvae10), so parity
is pinned for everything except that one third-party block ("parity unpinned" there).

Pinned by tests/golden/hyvae10_*.npz = outputs of the reference AutoencoderKLCausal3D.decode (oracle/gen_golden.py).
"""
import torch
import torch.nn.functional as F

from wan2gp_b200.synth import hyvae10_encoder_layout, hyvae10_layout
from .hyvae_oracle import _q, causal_conv3d_rep


def group_norm(x, w, b, groups, eps=1e-6):
    return F.group_norm(x[None], groups, w.float(), b.float(), eps)[0]


def resnet(sd, p, x, G, em):
    h = _q(F.silu(group_norm(x, sd[p + "norm1.weight"], sd[p + "norm1.bias"], G)), em)
    h = _q(causal_conv3d_rep(h, sd[p + "conv1.conv.weight"], sd[p + "conv1.conv.bias"], em), em)
    h = _q(F.silu(group_norm(h, sd[p + "norm2.weight"], sd[p + "norm2.bias"], G)), em)
    h = causal_conv3d_rep(h, sd[p + "conv2.conv.weight"], sd[p + "conv2.conv.bias"], em)
    if p + "conv_shortcut.conv.weight" in sd:
        x = _q(causal_conv3d_rep(x, sd[p + "conv_shortcut.conv.weight"], sd[p + "conv_shortcut.conv.bias"], em), em)
    return _q(h + x, em)                                      # output_scale_factor = 1 (vae.py:236, unet_causal_3d_blocks.py:790)


def attention(sd, p, x, G, causal, em):
    C, T, H, W = x.shape
    hn = _q(group_norm(x.reshape(C, T * H * W), sd[p + "group_norm.weight"], sd[p + "group_norm.bias"], G), em)
    tok = hn.t()
    q, k, v = (_q(tok @ _q(sd[p + n + ".weight"].float(), em).t() + sd[p + n + ".bias"].float(), em) for n in ("to_q", "to_k", "to_v"))
    s = (q @ k.t()) / (C ** 0.5)
    if causal:
        frame = torch.arange(T * H * W) // (H * W)
        s = s.masked_fill(frame[None, :] > frame[:, None], float("-inf"))
    pr = torch.exp(s - s.max(-1, keepdim=True).values)
    o = _q((_q(pr, em) @ v) / pr.sum(-1, keepdim=True), em)
    o = o @ _q(sd[p + "to_out.0.weight"].float(), em).t() + sd[p + "to_out.0.bias"].float()
    return _q(x + o.t().reshape(C, T, H, W), em)


def upsample(sd, p, x, up_t, up_s, em):
    """Nearest up-sampling; with time up-sampling the first frame is only up-sampled in space (:196-207)."""
    fs = 2 if up_s else 1
    if up_t:
        first = F.interpolate(x[None, :, :1], scale_factor=(1, fs, fs), mode="nearest")[0]
        x = torch.cat([first, F.interpolate(x[None, :, 1:], scale_factor=(2, fs, fs), mode="nearest")[0]], 1) if x.shape[1] > 1 else first
    else:
        x = F.interpolate(x[None], scale_factor=(1, fs, fs), mode="nearest")[0]
    return _q(causal_conv3d_rep(x, sd[p + "conv.conv.weight"], sd[p + "conv.conv.bias"], em), em)


def hyvae10_decode(sd, cfg, z, emulate_bf16=False):
    """z [zc,T,h,w] fp32 -> frames [3, 4(T-1)+1, 8h, 8w] fp32."""
    em, G = emulate_bf16, cfg["norm_num_groups"]
    z = z.float()
    z = F.conv3d(z[None], sd["post_quant_conv.weight"].float(), sd["post_quant_conv.bias"].float())[0]
    d = "decoder."
    h = _q(causal_conv3d_rep(_q(z, em), sd[d + "conv_in.conv.weight"], sd[d + "conv_in.conv.bias"], em), em)
    h = resnet(sd, d + "mid_block.resnets.0.", h, G, em)
    h = attention(sd, d + "mid_block.attentions.0.", h, G, cfg["mid_block_causal_attn"], em)
    h = resnet(sd, d + "mid_block.resnets.1.", h, G, em)
    blocks, _ = hyvae10_layout(cfg)
    for i, (rs, up) in enumerate(blocks):
        for j in range(len(rs)):
            h = resnet(sd, d + f"up_blocks.{i}.resnets.{j}.", h, G, em)
        if up is not None:
            h = upsample(sd, d + f"up_blocks.{i}.upsamplers.0.", h, up[0], up[1], em)
    h = _q(F.silu(group_norm(h, sd[d + "conv_norm_out.weight"], sd[d + "conv_norm_out.bias"], G)), em)
    return causal_conv3d_rep(h, sd[d + "conv_out.conv.weight"], sd[d + "conv_out.conv.bias"], em)


def downsample(sd, p, x, st_t, st_s, em):
    """DownsampleCausal3D (unet_causal_3d_blocks.py:226-298): replicate-padded causal 3x3x3 conv with stride (1|2, 2, 2)."""
    xp = F.pad(_q(x, em)[None], (1, 1, 1, 1, 2, 0), mode="replicate")
    y = F.conv3d(xp, _q(sd[p + "conv.conv.weight"].float(), em), sd[p + "conv.conv.bias"].float(), stride=(2 if st_t else 1, 2 if st_s else 1, 2 if st_s else 1))[0]
    return _q(y, em)


def hyvae10_encode(sd, cfg, x, emulate_bf16=False):
    """x [3, 1+4k, H, W] fp32 -> moments [2 zc, 1+k, H/8, W/8] fp32: EncoderCausal3D.forward (vae/vae.py:135-184) then quant_conv
    (autoencoder_kl_causal_3d.py:464-467, un-tiled branch)."""
    em, G = emulate_bf16, cfg["norm_num_groups"]
    e = "encoder."
    h = _q(causal_conv3d_rep(_q(x.float(), em), sd[e + "conv_in.conv.weight"], sd[e + "conv_in.conv.bias"], em), em)
    blocks, _ = hyvae10_encoder_layout(cfg)
    for i, (rs, down) in enumerate(blocks):
        for j in range(len(rs)):
            h = resnet(sd, e + f"down_blocks.{i}.resnets.{j}.", h, G, em)
        if down is not None:
            h = downsample(sd, e + f"down_blocks.{i}.downsamplers.0.", h, down[0], down[1], em)
    h = resnet(sd, e + "mid_block.resnets.0.", h, G, em)
    h = attention(sd, e + "mid_block.attentions.0.", h, G, cfg["mid_block_causal_attn"], em)
    h = resnet(sd, e + "mid_block.resnets.1.", h, G, em)
    h = _q(F.silu(group_norm(h, sd[e + "conv_norm_out.weight"], sd[e + "conv_norm_out.bias"], G)), em)
    h = causal_conv3d_rep(h, sd[e + "conv_out.conv.weight"], sd[e + "conv_out.conv.bias"], em)
    return F.conv3d(h[None], sd["quant_conv.weight"].float(), sd["quant_conv.bias"].float())[0]
