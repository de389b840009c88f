"""B200-native Hunyuan Video 1.5 VAE decode (hot-path row H6 of SURVEY.md section 8a): the `AutoencoderKLConv3D` surface the
pipeline uses (`.decode(z, return_dict=False)[0]`, `.enable_tiling()`, `.config.scaling_factor / .shift_factor`,
models/hyvideo/diffusion/pipelines/pipeline_hunyuan_video.py:1790-1812) over the reference Decoder
(models/hyvideo/vae/hunyuanvideo_15_vae.py:432-520).  Un-tiled by default (one B200 holds the whole clip); `enable_tiling()`
switches to the reference's temporal + spatial tiling with cross-faded seams (`_TiledDecode`), which is what its pipelines use.

Channels-last bf16 activations [T,H,W,C].  Replicate-padded causal convs (CausalConv3d :124-158) = one `pad_replicate` pass +
the tcgen05 implicit-GEMM conv over the padded tensor (TMA zero fill cannot replicate); 1x1x1 convs are plain GEMMs over
pixels; the mid block's single-head, frame-causal attention runs per query frame over keys of frames <= f.
"""
import math
import types

import torch

from .. import _lib, ops, synth
from ..wan.vae import _Conv, rms_silu

bf16, f32 = torch.bfloat16, torch.float32


def _s():
    return torch.cuda.current_stream().cuda_stream


PAD_SLICE_BYTES = 8 << 30            # budget for one padded, normalised time slice of a full-resolution activation


class _RepConv(_Conv):
    """3x3x3 causal conv with replicate padding: pad (2 frames in front, 1 pixel around) then a 'valid' conv."""

    def __call__(self, x, residual=None, out_mode=0):
        T, H, W, C = x.shape
        kt, kh, kw = self.k
        xp = torch.empty(T + kt - 1, H + kh - 1, W + kw - 1, C, device=x.device, dtype=bf16)
        _lib.call("b200_pad_replicate_cl", x.data_ptr(), xp.data_ptr(), T, H, W, C, kt - 1, kh // 2, kw // 2, _s())
        return self.prepadded(xp, T, H, W, residual, None, out_mode)

    def prepadded(self, xp, T, H, W, residual, out, out_mode):
        """'valid' conv over an already padded operand xp [T+kt-1, H+kh-1, W+kw-1, Cin]."""
        kt, kh, kw = self.k
        if out_mode == 2 and self.w_stack is not None and residual is None:
            return self.head(xp, T, H, W, True, out)
        if out is None:
            out = (torch.empty(self.cout, T, H, W, device=xp.device, dtype=f32) if out_mode == 2
                   else torch.empty(T, H, W, self.cout, device=xp.device, dtype=bf16))
        _lib.call("b200_conv3d_cl_prepadded", xp.data_ptr(), self.w.data_ptr(), self.b.data_ptr(),
                  0 if residual is None else residual.data_ptr(), out.data_ptr(), T, H, W, self.cin, self.cout, kt, kh, kw, out_mode, _s())
        return out


class _RMSNorm:
    """RMS_norm (hunyuanvideo_15_vae.py:107-122) with the norm_act_conv interface: no clip-wide statistics."""

    def __init__(self, gamma):
        self.g = gamma

    def stats(self, x):
        return None

    def apply(self, x, st, silu, t0=0, tc=None, pad=(0, 0, 0)):
        T, H, W, C = x.shape
        tc = T if tc is None else tc
        y = torch.empty(tc + pad[0], H + 2 * pad[1], W + 2 * pad[2], C, device=x.device, dtype=bf16)
        _lib.call("b200_rms_silu_pad_cl", x.data_ptr(), self.g.data_ptr(), y.data_ptr(), T, H, W, C, int(silu), t0, tc, pad[0], pad[1], pad[2], _s())
        return y


def norm_act_conv(x, norm, conv, residual=None, out_mode=0, out=None):
    """conv(replicate_pad(silu(norm(x)))) [+ residual].  norm -> SiLU -> pad is ONE pass that writes the padded operand of the
    tcgen05 implicit-GEMM conv, in time slices of <= PAD_SLICE_BYTES so the padded copy stays a few GB at 720p x 129 frames
    (a full-resolution activation is 30-61 GB).  `out` may alias `residual`: each epilogue thread reads its residual chunk
    before storing the same chunk."""
    T, H, W, C = x.shape
    kt, kh, kw = conv.k
    st = norm.stats(x)
    frame_bytes = (H + kh - 1) * (W + kw - 1) * C * 2
    tc_max = max(1, min(T, PAD_SLICE_BYTES // frame_bytes - (kt - 1)))
    if out is None:
        out = (torch.empty(conv.cout, T, H, W, device=x.device, dtype=f32) if out_mode == 2
               else torch.empty(T, H, W, conv.cout, device=x.device, dtype=f32 if out_mode == 3 else bf16))
    for t0 in range(0, T, tc_max):
        tc = min(tc_max, T - t0)
        xp = norm.apply(x, st, True, t0, tc, (kt - 1, kh // 2, kw // 2))
        if out_mode == 2 and tc != T:                       # planar fp32 head: the channel stride is the slice's, so copy it in
            out[:, t0:t0 + tc].copy_(conv.prepadded(xp, tc, H, W, None, None, 2))
        else:
            conv.prepadded(xp, tc, H, W, None if residual is None else residual[t0:t0 + tc], out if out_mode == 2 else out[t0:t0 + tc], out_mode)
        del xp
    return out


class HYVAEDecoder(torch.nn.Module):
    def __init__(self, cfg, device="cuda"):
        super().__init__()
        self.cfg, self.device = dict(cfg), torch.device(device)
        self._ready = False

    def load_state_dict(self, sd, strict=True, assign=False):
        dev = self.device
        gam = lambda k: sd[k].detach().to(dev, f32).reshape(-1).contiguous()                      # noqa: E731
        rms = lambda k: _RMSNorm(gam(k))                                                           # noqa: E731
        rc = lambda p: _RepConv(sd[p + ".weight"], sd[p + ".bias"], dev)                           # noqa: E731
        lin = lambda p: (sd[p + ".weight"].detach().to(dev, bf16).reshape(sd[p + ".weight"].shape[0], -1).contiguous(),  # noqa: E731
                         sd[p + ".bias"].detach().to(dev, f32).contiguous())

        def res(p):
            d = {"g1": rms(p + "norm1.gamma"), "c1": rc(p + "conv1.conv"), "g2": rms(p + "norm2.gamma"), "c2": rc(p + "conv2.conv")}
            if p + "nin_shortcut.weight" in sd:
                d["nin"] = lin(p + "nin_shortcut")
            return d
        self.conv_in = rc("conv_in.conv")
        self.mid1, self.mid2 = res("mid.block_1."), res("mid.block_2.")
        a = "mid.attn_1."
        self.attn = {"g": gam(a + "norm.gamma"),
                     "wqkv": torch.cat([lin(a + n)[0] for n in "qkv"], 0).contiguous(), "bqkv": torch.cat([lin(a + n)[1] for n in "qkv"], 0).contiguous(),
                     "proj": lin(a + "proj_out")}
        self.levels = []
        levels, _ = synth.hyvae_layout(self.cfg)
        for i, (blocks, up) in enumerate(levels):
            self.levels.append(([res(f"up.{i}.block.{j}.") for j in range(len(blocks))],
                                None if up is None else (rc(f"up.{i}.upsample.conv.conv"), up[1], up[2])))
        self.g_out = rms("norm_out.gamma")
        self.conv_out = rc("conv_out.conv")
        self._ready = True
        return torch.nn.modules.module._IncompatibleKeys([], [])

    @staticmethod
    def _res(d, hold):
        """ResnetBlock.forward (:217-250).  `hold` is a one-element list that is CONSUMED: the block input is freed (or
        overwritten in place by the output) as soon as the shortcut exists."""
        x = hold.pop()
        T, H, W, C = x.shape
        h = norm_act_conv(x, d["g1"], d["c1"])
        sc = x
        if "nin" in d:                                                       # 1x1x1 conv == GEMM over pixels
            sc = ops.gemm(x.reshape(-1, C), d["nin"][0], bias=d["nin"][1]).reshape(T, H, W, -1)
        del x
        return norm_act_conv(h, d["g2"], d["c2"], residual=sc, out=sc)       # h += x fused in the conv epilogue, in place

    def _attn(self, x):
        T, H, W, C = x.shape
        a, N = self.attn, H * W
        xn = rms_silu(x, a["g"], silu=False).reshape(T * N, C)
        buf = torch.zeros(T * N + 8, 3 * C, device=x.device, dtype=bf16)
        qkv = ops.gemm(xn, a["wqkv"], out=buf[:T * N], bias=a["bqkv"])
        npad = (T * N + 63) // 64 * 64
        ws = torch.empty(N * npad * 6, device=x.device, dtype=torch.uint8)
        o = torch.empty(T * N, C, device=x.device, dtype=bf16)
        _lib.call("b200_attention_1head", qkv.data_ptr(), o.data_ptr(), ws.data_ptr(), ws.numel(), T, N, C, float(C) ** -0.5, 1, _s())
        return ops.gemm(o, a["proj"][0], bias=a["proj"][1], residual=x.reshape(T * N, C)).reshape(T, H, W, C)

    @staticmethod
    def _up(up, x):
        conv, cout, temporal = up
        T, H, W, Ci = x.shape
        h = conv(x)
        out = torch.empty((2 * T - 1) if temporal else T, 2 * H, 2 * W, cout, device=x.device, dtype=bf16)
        _lib.call("b200_hy_upsample_cl", h.data_ptr(), x.data_ptr(), out.data_ptr(), T, H, W, Ci, cout, int(temporal), _s())
        return out

    @torch.no_grad()
    def forward(self, z):
        """z [B, zc, T, h, w] -> frames fp32 [B, 3, ft*(T-1)+1, fs*h, fs*w] (Decoder.forward, :486-520)."""
        if not self._ready:
            raise RuntimeError("HYVAEDecoder: load_state_dict() must be called before decode")
        outs = []
        c0, zc = self.cfg["block_out_channels"][0], self.cfg["z_channels"]
        for zi in z:
            zi = zi.to(self.device, f32).contiguous()
            _, T, H, W = zi.shape
            P = T * H * W
            zcl = torch.empty(T, H, W, zc, device=self.device, dtype=bf16)
            zrep = torch.empty(T, H, W, c0, device=self.device, dtype=bf16)
            _lib.call("b200_planar_to_cl", zi.data_ptr(), zcl.data_ptr(), zc, P, 1, _s())
            _lib.call("b200_planar_to_cl", zi.data_ptr(), zrep.data_ptr(), zc, P, c0 // zc, _s())
            hold = [self.conv_in(zcl, residual=zrep)]                          # conv_in(z) + z.repeat_interleave (:489-490)
            del zcl, zrep
            hold.append(self._res(self.mid1, hold))
            hold.append(self._attn(hold.pop()))
            hold.append(self._res(self.mid2, hold))
            for blocks, up in self.levels:
                for d in blocks:
                    hold.append(self._res(d, hold))
                if up is not None:
                    hold.append(self._up(up, hold.pop()))
            outs.append(norm_act_conv(hold.pop(), self.g_out, self.conv_out, out_mode=2))
        return torch.stack(outs, 0)


class HYVAEEncoder(torch.nn.Module):
    """Hunyuan Video 1.5 VAE Encoder.forward (hunyuanvideo_15_vae.py:395-430), un-tiled: x [B,3,1+4k,H,W] -> posterior moments
    [B, 2 zc, 1+k, H/fs, W/fs] fp32 (mean | logvar).  Same kernels as the decoder; Downsample (:253-296) = causal conv + one
    shuffle / group-mean pass (`b200_hy_downsample_cl`), no strided convolution anywhere."""

    def __init__(self, cfg, device="cuda"):
        super().__init__()
        self.cfg, self.device = dict(cfg), torch.device(device)
        self._ready = False

    def load_state_dict(self, sd, strict=True, assign=False):
        dev = self.device
        gam = lambda k: sd[k].detach().to(dev, f32).reshape(-1).contiguous()                      # noqa: E731
        rms = lambda k: _RMSNorm(gam(k))                                                           # noqa: E731
        rc = lambda p: _RepConv(sd[p + ".weight"], sd[p + ".bias"], dev)                           # noqa: E731
        lin = lambda p: (sd[p + ".weight"].detach().to(dev, bf16).reshape(sd[p + ".weight"].shape[0], -1).contiguous(),  # noqa: E731
                         sd[p + ".bias"].detach().to(dev, f32).contiguous())

        def res(p):
            d = {"g1": rms(p + "norm1.gamma"), "c1": rc(p + "conv1.conv"), "g2": rms(p + "norm2.gamma"), "c2": rc(p + "conv2.conv")}
            if p + "nin_shortcut.weight" in sd:
                d["nin"] = lin(p + "nin_shortcut")
            return d
        w_in = sd["conv_in.conv.weight"].detach().to(dev, f32)
        self.cin = w_in.shape[1]
        self.conv_in = _RepConv(torch.cat([w_in, w_in.new_zeros(w_in.shape[0], 8 - self.cin, *w_in.shape[2:])], 1), sd["conv_in.conv.bias"], dev)
        levels, self.c_mid = synth.hyvae_encoder_layout(self.cfg)
        self.levels = []
        for i, (blocks, down) in enumerate(levels):
            dn = None
            if down is not None:                   # conv output channels padded to a multiple of 16 (zero weights); the shuffle reads the real ones
                w, b = sd[f"down.{i}.downsample.conv.conv.weight"].detach().to(dev, f32), sd[f"down.{i}.downsample.conv.conv.bias"].detach().to(dev, f32)
                padc = (-w.shape[0]) % 16
                dn = (_RepConv(torch.cat([w, w.new_zeros(padc, *w.shape[1:])], 0), torch.cat([b, b.new_zeros(padc)]), dev), down[1], down[2])
            self.levels.append(([res(f"down.{i}.block.{j}.") for j in range(len(blocks))], dn))
        self.mid1, self.mid2 = res("mid.block_1."), res("mid.block_2.")
        a = "mid.attn_1."
        self.attn = {"g": gam(a + "norm.gamma"),
                     "wqkv": torch.cat([lin(a + n)[0] for n in "qkv"], 0).contiguous(), "bqkv": torch.cat([lin(a + n)[1] for n in "qkv"], 0).contiguous(),
                     "proj": lin(a + "proj_out")}
        self.g_out = rms("norm_out.gamma")
        self.conv_out = rc("conv_out.conv")
        self._ready = True
        return torch.nn.modules.module._IncompatibleKeys([], [])

    _res = staticmethod(HYVAEDecoder._res)
    _attn = HYVAEDecoder._attn

    @torch.no_grad()
    def forward(self, x):
        if not self._ready:
            raise RuntimeError("HYVAEEncoder: load_state_dict() must be called before encode")
        outs = []
        zc2 = 2 * self.cfg["z_channels"]
        for xi in x:
            xi = xi.to(self.device, f32).contiguous()
            C, T, H, W = xi.shape
            fs, ft = self.cfg["ffactor_spatial"], self.cfg["ffactor_temporal"]
            if C != self.cin or (T - 1) % ft or H % fs or W % fs:
                raise ValueError(f"Hunyuan VAE encode: expected [{self.cin}, 1+{ft}k, {fs}m, {fs}n] frames, got {tuple(xi.shape)}")
            xcl = torch.empty(T, H, W, 8, device=self.device, dtype=bf16)
            _lib.call("b200_planar_to_cl_pad", xi.data_ptr(), xcl.data_ptr(), C, T * H * W, 8, _s())
            hold = [self.conv_in(xcl)]
            del xcl
            for blocks, down in self.levels:
                for d in blocks:
                    hold.append(self._res(d, hold))
                if down is not None:
                    conv, cout, temporal = down
                    h = hold.pop()
                    t, hh, ww, ci = h.shape
                    hc = conv(h)
                    out = torch.empty(1 + (t - 1) // 2 if temporal else t, hh // 2, ww // 2, cout, device=self.device, dtype=bf16)
                    _lib.call("b200_hy_downsample_cl", hc.data_ptr(), hc.shape[-1], h.data_ptr(), out.data_ptr(), t, hh, ww, ci, cout, int(temporal), _s())
                    del h, hc
                    hold.append(out)
            hold.append(self._res(self.mid1, hold))
            hold.append(self._attn(hold.pop()))
            hold.append(self._res(self.mid2, hold))
            h = hold.pop()
            t, hh, ww, c = h.shape
            sc = torch.empty(t, hh, ww, zc2, device=self.device, dtype=bf16)                    # channel-group mean shortcut (:424-425)
            _lib.call("b200_group_mean_cl", h.data_ptr(), sc.data_ptr(), t * hh * ww, c, c // zc2, _s())
            mom = norm_act_conv(h, self.g_out, self.conv_out, residual=sc, out_mode=3)           # fp32 [T,h,w,2zc]
            outs.append(mom.permute(3, 0, 1, 2).contiguous())
        return torch.stack(outs, 0)


class _Posterior:
    """diffusers' DiagonalGaussianDistribution surface the callers use (`.mode()`, `.sample()`, `.mean`, `.logvar`)."""

    def __init__(self, moments):
        self.mean, self.logvar = moments.chunk(2, dim=1)
        self.logvar = self.logvar.clamp(-30.0, 20.0)

    def mode(self):
        return self.mean

    def sample(self, generator=None):
        eps = torch.randn(self.mean.shape, generator=generator, device=self.mean.device if generator is None else generator.device, dtype=f32)
        return self.mean + torch.exp(0.5 * self.logvar) * eps.to(self.mean.device)


class _TiledDecode:
    """The tiling switches and dispatch both Hunyuan VAE classes share (hunyuanvideo_15_vae.py:582-625, 806-864, 889-896;
    autoencoder_kl_causal_3d.py:297-330, 474-482, 638-855).  The pipelines call enable_tiling() before every decode (hunyuan.py:772,
    pipeline_hunyuan_video.py:695) and get_VAE_tile_size() picks 256 px x 64 frames even on a large GPU, so the reference's ACTUAL output is
    the tiled + cross-faded one; this reproduces it tile for tile (each tile = one whole-clip decode on the B200 kernels, seams blended
    by b200_blend_edge_f32).  disable_tiling() gives the un-tiled whole-clip decode (one pass, no seams, ~2x fewer FLOPs)."""

    tile_overlap_factor = 0.25
    use_spatial_tiling = use_temporal_tiling = False

    def _init_tiles(self, sample_size, sample_tsize, lat_size, lat_tsize):
        self.tile_sample_min_size, self.tile_sample_min_tsize = sample_size, sample_tsize
        self.tile_latent_min_size, self.tile_latent_min_tsize = lat_size, lat_tsize

    def enable_spatial_tiling(self, use_tiling=True):
        self.use_spatial_tiling = use_tiling

    def enable_temporal_tiling(self, use_tiling=True):
        self.use_temporal_tiling = use_tiling

    def enable_tiling(self, use_tiling=True):
        self.enable_spatial_tiling(use_tiling), self.enable_temporal_tiling(use_tiling)

    def disable_spatial_tiling(self):
        self.enable_spatial_tiling(False)

    def disable_temporal_tiling(self):
        self.enable_temporal_tiling(False)

    def disable_tiling(self):
        self.enable_tiling(False)

    def enable_slicing(self):
        return None                              # batch items are decoded one at a time anyway

    disable_slicing = enable_slicing

    def _tiled_clip(self, x, fn, in_s, in_t, out_s, out_t):
        """x [C,T,h,w] -> fn applied tile by tile with the reference's dispatch.  Decode: in = latent tile sizes, out = sample tile sizes
        (temporal_tiled_decode / spatial_tiled_decode); encode: the other way round (temporal_tiled_encode / spatial_tiled_encode).
        Strides come from the INPUT tile, blend extents and crop limits from the OUTPUT tile."""
        from ..wan.vae import _lib as lib, spatial_tiles
        ov = self.tile_overlap_factor
        one = lambda t: fn(t[None])[0]                                                                # noqa: E731

        def sp(t):
            if self.use_spatial_tiling and (t.shape[-1] > in_s or t.shape[-2] > in_s):
                blend = int(out_s * ov)
                return spatial_tiles(t, in_s, int(in_s * (1 - ov)), one, blend, out_s - blend)
            return one(t.contiguous())
        if not (self.use_temporal_tiling and x.shape[1] > in_t):
            return sp(x)
        stride, blend = int(in_t * (1 - ov)), int(out_t * ov)
        if not 0 < stride < in_t:
            raise ValueError("temporal tile stride must be in (0, tile size)")
        t_limit = out_t - blend
        row = []
        for i in range(0, x.shape[1], stride):
            d = sp(x[:, i:i + in_t + 1])
            row.append(d[:, 1:].contiguous() if i > 0 else d)
        out = []
        for i, t in enumerate(row):
            if i > 0 and t.shape[1] > 0:
                a = row[i - 1]                     # blend_t: frames are the "rows" of a [C, F, H*W] tile
                lib.call("b200_blend_edge_f32", a.data_ptr(), t.data_ptr(), a.shape[0], a.shape[1], a.shape[2] * a.shape[3], t.shape[1],
                         t.shape[2] * t.shape[3], blend, 1, _s())
            out.append(t[:, :t_limit + (1 if i == 0 else 0)])
        return torch.cat(out, 1)

    def _decode_batch(self, z):
        if not (self.use_spatial_tiling or self.use_temporal_tiling):
            return self.decoder(z)
        args = (self.tile_latent_min_size, self.tile_latent_min_tsize, self.tile_sample_min_size, self.tile_sample_min_tsize)
        return torch.stack([self._tiled_clip(zi.to(self.decoder.device, f32), self.decoder, *args) for zi in z], 0)

    def _encode_batch(self, x):
        if not (self.use_spatial_tiling or self.use_temporal_tiling):
            return self.encoder(x)
        args = (self.tile_sample_min_size, self.tile_sample_min_tsize, self.tile_latent_min_size, self.tile_latent_min_tsize)
        return torch.stack([self._tiled_clip(xi.to(self.encoder.device, f32), self.encoder, *args) for xi in x], 0)


class AutoencoderKLConv3D(_TiledDecode, torch.nn.Module):
    """Decode surface of models/hyvideo/vae/hunyuanvideo_15_vae.py::AutoencoderKLConv3D (:523-907)."""

    def __init__(self, in_channels=3, out_channels=3, latent_channels=32, block_out_channels=(128, 256, 512, 1024, 1024),
                 layers_per_block=2, ffactor_spatial=16, ffactor_temporal=4, sample_size=256, sample_tsize=64, scaling_factor=None,
                 shift_factor=None, device="cuda", **unused):
        super().__init__()
        self.ffactor_spatial, self.ffactor_temporal = ffactor_spatial, ffactor_temporal
        self.scaling_factor, self.shift_factor = scaling_factor, shift_factor
        self.config = types.SimpleNamespace(scaling_factor=scaling_factor, shift_factor=shift_factor, latent_channels=latent_channels)
        cfg = dict(z_channels=latent_channels, out_channels=out_channels, block_out_channels=list(reversed(list(block_out_channels))),
                   num_res_blocks=layers_per_block, ffactor_spatial=ffactor_spatial, ffactor_temporal=ffactor_temporal)
        self.decoder = HYVAEDecoder(cfg, device)
        self.encoder = HYVAEEncoder(cfg, device)
        self._init_tiles(sample_size, sample_tsize, sample_size // ffactor_spatial, sample_tsize // ffactor_temporal)     # :568-575

    def load_state_dict(self, sd, strict=True, assign=False):
        enc = {k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")}
        if enc:                                    # decode-only checkpoints carry no encoder
            self.encoder.load_state_dict(enc)
        return self.decoder.load_state_dict({k[len("decoder."):]: v for k, v in sd.items() if k.startswith("decoder.")})

    def set_tile_sample_min_size(self, sample_size, tile_overlap_factor=0.25, sample_tsize=None):
        """hunyuanvideo_15_vae.py:582-594."""
        self.tile_sample_min_size, self.tile_latent_min_size = sample_size, sample_size // self.ffactor_spatial
        self.tile_overlap_factor = tile_overlap_factor
        if sample_tsize is not None:
            self.tile_sample_min_tsize, self.tile_latent_min_tsize = sample_tsize, sample_tsize // self.ffactor_temporal

    def decode(self, z, return_dict=True, generator=None):
        out = self._decode_batch(z)
        return types.SimpleNamespace(sample=out) if return_dict else (out,)

    def encode(self, x, return_dict=True):
        """AutoencoderKLConv3D.encode (:866-887), tiling off: posterior over the encoder's moments."""
        post = _Posterior(self._encode_batch(x))
        return types.SimpleNamespace(latent_dist=post) if return_dict else (post,)
