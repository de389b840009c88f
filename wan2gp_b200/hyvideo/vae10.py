"""B200-native HunyuanVideo 1.0 VAE decode (hot-path row H5 of SURVEY.md section 8a): the `AutoencoderKLCausal3D` surface the
pipeline uses (`.decode(z, return_dict=False)[0]`, `.enable_tiling()`, `.config.scaling_factor`,
models/hyvideo/diffusion/pipelines/pipeline_hunyuan_video.py:1790-1812) over post_quant_conv + DecoderCausal3D
(models/hyvideo/vae/autoencoder_kl_causal_3d.py:474-493, vae/vae.py:186-365, vae/unet_causal_3d_blocks.py).  Un-tiled by default
(one B200 holds the whole clip); `enable_tiling()` switches to the reference's temporal + spatial tiling with cross-faded seams
(hyvideo/vae.py::_TiledDecode), which is what its pipelines use.

Channels-last bf16 activations [T,H,W,C].
  * GroupNorm (clip-wide statistics) -> SiLU -> replicate pad is ONE elementwise pass after a statistics pass: it writes the
    padded tensor the tcgen05 implicit-GEMM conv reads, in time slices so the padded copy stays a few GB even at 720p x 129f.
  * nearest up-sampling (first frame spatial-only, :196-207) + CausalConv3d is evaluated on the LOW-resolution tensor as 4
    (x1,2,2) or 8 (x2,2,2) phase convs with pre-summed taps (2x2 / 2x2x2 instead of 3x3(x3)): 2.25x / 3.4x fewer FLOPs and the
    up-sampled tensor (61 GB at 720p x 129f x 256 ch) is never materialised.
  * post_quant_conv (1x1x1) is folded into conv_in's weights at load time (a pointwise map commutes with replicate padding).
  * mid block: single-head attention over all T*H*W tokens with the frame-causal mask (:21-30), as for the 1.5 VAE.
"""
import types

import torch

from .. import _lib, ops, synth
from .vae import _Posterior, _RepConv, _TiledDecode, norm_act_conv

bf16, f32 = torch.bfloat16, torch.float32


def _s():
    return torch.cuda.current_stream().cuda_stream


class _GroupNorm:
    def __init__(self, w, b, groups, device, eps=1e-6):
        self.g, self.b = w.detach().to(device, f32).contiguous(), b.detach().to(device, f32).contiguous()
        self.groups, self.eps = groups, eps
        self.ws = torch.empty(148 * 8 * groups * 2, device=device, dtype=f32)        # B200_GROUP_STATS_WS_BYTES

    def stats(self, x):
        T, H, W, C = x.shape
        st = torch.empty(2 * C, device=x.device, dtype=f32)                  # per-channel (scale, shift) of the normalisation
        _lib.call("b200_group_stats_cl", x.data_ptr(), self.g.data_ptr(), self.b.data_ptr(), st.data_ptr(), self.ws.data_ptr(),
                  T * H * W, C, self.groups, self.eps, _s())
        return st

    def apply(self, x, st, silu, t0=0, tc=None, pad=(0, 0, 0)):
        T, H, W, C = x.shape
        tc = T if tc is None else tc
        y = torch.empty(tc + pad[0], H + 2 * pad[1], W + 2 * pad[2], C, device=x.device, dtype=bf16)
        _lib.call("b200_group_norm_apply_cl", x.data_ptr(), st.data_ptr(), y.data_ptr(), T, H, W, C, int(silu), t0, tc,
                  pad[0], pad[1], pad[2], _s())
        return y


class _UpConvNearest:
    """UpsampleCausal3D (unet_causal_3d_blocks.py:96-224): nearest x(ft,2,2) then a replicate-padded 3x3x3 causal conv, folded
    into phase convs on the low-resolution tensor.  With u = up-sampled index, source index s(u) = u // 2 in space; in time
    s(0) = 0 and s(u) = 1 + (u-1)//2 (first frame not repeated).  Output parity 0 reads sources (i-1, i) with taps
    {k0 | k1+k2}; parity 1 reads (i, i+1) with {k0+k1 | k2} -- in time: even frame 2m reads (m-1, m), odd frame 2m+1 reads
    (m, m+1).  Replicate padding of the up-sampled tensor == replicate padding of the source."""

    def __init__(self, w, b, up_t, device, dtype=bf16):
        w = w.detach().to(device, f32)                                  # [Co, Ci, 3, 3, 3]
        self.cout, self.cin, self.up_t = w.shape[0], w.shape[1], up_t
        self.b = b.detach().to(device, f32).contiguous()

        def fold(t, dim, parity):
            k0, k1, k2 = t.select(dim, 0), t.select(dim, 1), t.select(dim, 2)
            return torch.stack([k0, k1 + k2] if parity == 0 else [k0 + k1, k2], dim)
        self.w = {}
        for pt in ((0, 1) if up_t else (None,)):
            wt = w if pt is None else fold(w, 2, pt)
            for py in (0, 1):
                for px in (0, 1):
                    wp = fold(fold(wt, 3, py), 4, px)                    # [Co, Ci, kt', 2, 2]
                    self.w[(pt, py, px)] = wp.permute(0, 2, 3, 4, 1).reshape(self.cout, -1, self.cin).to(dtype).contiguous()

    def __call__(self, x):
        T, H, W, C = x.shape
        ptf = 1 if self.up_t else 2                                      # frames of front padding the phase convs need
        xp = torch.empty(T + ptf, H + 2, W + 2, C, device=x.device, dtype=bf16)
        _lib.call("b200_pad_replicate_cl", x.data_ptr(), xp.data_ptr(), T, H, W, C, ptf, 1, 1, _s())
        To = 2 * T - 1 if self.up_t else T
        Co = self.cout
        out = torch.empty(To, 2 * H, 2 * W, Co, device=x.device, dtype=bf16)
        for (pt, py, px), w in self.w.items():
            if pt is None:
                n_t, off_t, o_t, st_t, kt = T, 0, 0, 4 * H * W * Co, 3
            else:
                n_t, off_t, o_t, st_t, kt = (T, 0, 0, 8 * H * W * Co, 2) if pt == 0 else (T - 1, 1, 1, 8 * H * W * Co, 2)
            if n_t <= 0:
                continue
            base = out.data_ptr() + 2 * ((o_t * 2 * H + py) * 2 * W + px) * Co
            _lib.call("b200_conv3d_cl_view", xp.data_ptr(), T + ptf, H + 2, W + 2, off_t, py, px, w.data_ptr(), self.b.data_ptr(), 0, base,
                      n_t, H, W, self.cin, Co, kt, 2, 2, st_t, 4 * W * Co, 2 * Co, _s())
        return out


class _DownConvRep:
    """DownsampleCausal3D (unet_causal_3d_blocks.py:226-298): replicate-padded causal 3x3x3 conv with stride (1|2, 2, 2), on the
    stride-1 conv kernels.  Space: the PADDED tensor goes through space-to-depth ([.., (H+2)/2, (W+2)/2, 4C]) and the 3x3 taps become
    2x2 taps (kernel row a = 2*da + p, row 3 gets zero weight).  Time (stride 2): output t reads padded frames 2t, 2t+1, 2t+2 -- with
    the frames viewed as PAIRS stacked along H, taps (w0, w2) are a 2-frame conv over the even window and w1 a 1-frame conv over the
    odd window, accumulated through the conv's residual input."""

    def __init__(self, w, b, st_t, device, dtype=bf16):
        co, ci = w.shape[:2]
        w = w.detach().to(device, f32)                                      # [Co, Ci, 3, 3, 3]
        w2 = torch.zeros(co, 3, 2, 2, 2, 2, ci, device=device, dtype=f32)   # [Co, dt, da, db, p, q, Ci]
        for da in range(2):
            for p in range(2):
                for db in range(2):
                    for q in range(2):
                        if 2 * da + p < 3 and 2 * db + q < 3:
                            w2[:, :, da, db, p, q] = w[:, :, :, 2 * da + p, 2 * db + q].permute(0, 2, 1)
        w2 = w2.reshape(co, 3, 4, 4 * ci)                                    # [Co][dt][(da,db)][(p,q,c)]
        self.cout, self.cin, self.st_t = co, ci, st_t
        self.b = b.detach().to(device, f32).contiguous()
        if st_t:
            self.w_even = w2[:, [0, 2]].reshape(co, 8, 4 * ci).to(dtype).contiguous()
            self.w_odd = w2[:, 1].reshape(co, 4, 4 * ci).to(dtype).contiguous()
        else:
            self.w_all = w2.reshape(co, 12, 4 * ci).to(dtype).contiguous()

    def __call__(self, x):
        T, H, W, C = x.shape
        if H % 2 or W % 2:
            raise ValueError("Hunyuan VAE encode: frame height and width must be even at every level")
        xp = torch.empty(T + 2, H + 2, W + 2, C, device=x.device, dtype=bf16)
        _lib.call("b200_pad_replicate_cl", x.data_ptr(), xp.data_ptr(), T, H, W, C, 2, 1, 1, _s())
        hs, ws, h, w, co = (H + 2) // 2, (W + 2) // 2, H // 2, W // 2, self.cout
        s2d = torch.empty(T + 2 + (T + 2) % 2, hs, ws, 4 * C, device=x.device, dtype=bf16)          # even frame count for the pair view
        _lib.call("b200_space_to_depth_cl", xp.data_ptr(), s2d.data_ptr(), T + 2, H + 2, W + 2, C, _s())
        del xp
        st = (h * w * co, w * co, co)
        if not self.st_t:
            out = torch.empty(T, h, w, co, device=x.device, dtype=bf16)
            _lib.call("b200_conv3d_cl_view", s2d.data_ptr(), T + 2, hs, ws, 0, 0, 0, self.w_all.data_ptr(), self.b.data_ptr(), 0, out.data_ptr(),
                      T, h, w, 4 * C, co, 3, 2, 2, *st, _s())
            return out
        to, ns = (T - 1) // 2 + 1, s2d.shape[0] // 2
        out = torch.empty(to, h, w, co, device=x.device, dtype=bf16)
        _lib.call("b200_conv3d_cl_view", s2d.data_ptr(), ns, 2 * hs, ws, 0, 0, 0, self.w_even.data_ptr(), self.b.data_ptr(), 0, out.data_ptr(),
                  to, h, w, 4 * C, co, 2, 2, 2, *st, _s())
        _lib.call("b200_conv3d_cl_view", s2d.data_ptr(), ns, 2 * hs, ws, 0, hs, 0, self.w_odd.data_ptr(), 0, out.data_ptr(), out.data_ptr(),
                  to, h, w, 4 * C, co, 1, 2, 2, *st, _s())
        return out


class HYVAE10Encoder(torch.nn.Module):
    """EncoderCausal3D.forward + quant_conv (vae/vae.py:135-184, autoencoder_kl_causal_3d.py:464-467), un-tiled:
    x [B,3,1+4k,H,W] -> posterior moments [B, 2 zc, 1+k, H/8, W/8] fp32.  quant_conv (1x1x1) is folded into conv_out."""

    def __init__(self, cfg, device="cuda"):
        super().__init__()
        self.cfg, self.device = dict(cfg), torch.device(device)
        self._ready = False

    def load_state_dict(self, sd, strict=True, assign=False):
        dev, G = self.device, self.cfg["norm_num_groups"]
        gn = lambda p: _GroupNorm(sd[p + ".weight"], sd[p + ".bias"], G, dev)                        # noqa: E731
        rc = lambda p: _RepConv(sd[p + ".weight"], sd[p + ".bias"], dev)                             # noqa: E731
        lin = lambda p: (sd[p + ".weight"].detach().to(dev, bf16).reshape(sd[p + ".weight"].shape[0], -1).contiguous(),  # noqa: E731
                         sd[p + ".bias"].detach().to(dev, f32).contiguous())

        def res(p):
            d = {"n1": gn(p + "norm1"), "c1": rc(p + "conv1.conv"), "n2": gn(p + "norm2"), "c2": rc(p + "conv2.conv")}
            if p + "conv_shortcut.conv.weight" in sd:
                d["sc"] = lin(p + "conv_shortcut.conv")
            return d
        e = "encoder."
        w_in = sd[e + "conv_in.conv.weight"].detach().to(dev, f32)
        self.cin = w_in.shape[1]
        self.conv_in = _RepConv(torch.cat([w_in, w_in.new_zeros(w_in.shape[0], 8 - self.cin, *w_in.shape[2:])], 1), sd[e + "conv_in.conv.bias"], dev)
        blocks, _ = synth.hyvae10_encoder_layout(self.cfg)
        self.blocks = []
        for i, (rs, down) in enumerate(blocks):
            dn = None
            if down is not None:
                if not down[1]:
                    raise NotImplementedError("time-only down-sampling block (not produced by the 884 layout)")
                p = e + f"down_blocks.{i}.downsamplers.0.conv.conv"
                dn = _DownConvRep(sd[p + ".weight"], sd[p + ".bias"], down[0], dev)
            self.blocks.append(([res(e + f"down_blocks.{i}.resnets.{j}.") for j in range(len(rs))], dn))
        self.mid1, self.mid2 = res(e + "mid_block.resnets.0."), res(e + "mid_block.resnets.1.")
        a = e + "mid_block.attentions.0."
        self.attn = {"gn": gn(a + "group_norm"),
                     "wqkv": torch.cat([lin(a + n)[0] for n in ("to_q", "to_k", "to_v")], 0).contiguous(),
                     "bqkv": torch.cat([lin(a + n)[1] for n in ("to_q", "to_k", "to_v")], 0).contiguous(), "proj": lin(a + "to_out.0")}
        self.norm_out = gn(e + "conv_norm_out")
        # quant_conv o conv_out: W' = Wq Wo, b' = Wq bo + bq
        wo, bo = sd[e + "conv_out.conv.weight"].detach().to(dev, f32), sd[e + "conv_out.conv.bias"].detach().to(dev, f32)
        wq, bq = sd["quant_conv.weight"].detach().to(dev, f32), sd["quant_conv.bias"].detach().to(dev, f32)
        wq = wq.reshape(wq.shape[0], wq.shape[1])
        self.conv_out = _RepConv(torch.einsum("om,mctyx->octyx", wq, wo), wq @ bo + bq, dev)
        self._ready = True
        return torch.nn.modules.module._IncompatibleKeys([], [])

    @torch.no_grad()
    def forward(self, x):
        if not self._ready:
            raise RuntimeError("HYVAE10Encoder: load_state_dict() must be called before encode")
        outs = []
        for xi in x:
            xi = xi.to(self.device, f32).contiguous()
            C, T, H, W = xi.shape
            if C != self.cin or (T - 1) % 4 or H % 8 or W % 8:
                raise ValueError(f"Hunyuan VAE encode: expected [{self.cin}, 1+4k, 8m, 8n] frames, got {tuple(xi.shape)}")
            xcl = torch.empty(T, H, W, 8, device=self.device, dtype=bf16)
            _lib.call("b200_planar_to_cl_pad", xi.data_ptr(), xcl.data_ptr(), C, T * H * W, 8, _s())
            hold = [self.conv_in(xcl)]
            del xcl
            for rs, dn in self.blocks:
                for d in rs:
                    hold.append(self._res(d, hold))
                if dn is not None:
                    hold.append(dn(hold.pop()))
            hold.append(self._res(self.mid1, hold))
            hold.append(self._attn(hold.pop()))
            hold.append(self._res(self.mid2, hold))
            mom = norm_act_conv(hold.pop(), self.norm_out, self.conv_out, out_mode=3)               # fp32 [T,h,w,2zc]
            outs.append(mom.permute(3, 0, 1, 2).contiguous())
        return torch.stack(outs, 0)


class HYVAE10Decoder(torch.nn.Module):
    """post_quant_conv + DecoderCausal3D.forward (vae/vae.py:300-365)."""

    def __init__(self, cfg, device="cuda"):
        super().__init__()
        self.cfg, self.device = dict(cfg), torch.device(device)
        self._ready = False

    def load_state_dict(self, sd, strict=True, assign=False):
        dev, G = self.device, self.cfg["norm_num_groups"]
        gn = lambda p: _GroupNorm(sd[p + ".weight"], sd[p + ".bias"], G, dev)                        # noqa: E731
        rc = lambda p: _RepConv(sd[p + ".weight"], sd[p + ".bias"], dev)                               # noqa: E731
        lin = lambda p: (sd[p + ".weight"].detach().to(dev, bf16).reshape(sd[p + ".weight"].shape[0], -1).contiguous(),  # noqa: E731
                         sd[p + ".bias"].detach().to(dev, f32).contiguous())

        def res(p):
            d = {"n1": gn(p + "norm1"), "c1": rc(p + "conv1.conv"), "n2": gn(p + "norm2"), "c2": rc(p + "conv2.conv")}
            if p + "conv_shortcut.conv.weight" in sd:
                d["sc"] = lin(p + "conv_shortcut.conv")
            return d
        # conv_in o post_quant_conv: W'[o,i,tap] = sum_m W[o,m,tap] P[m,i];  b' = b + sum_{m,tap} W[o,m,tap] pb[m]
        d = "decoder."
        wq, bq = sd["post_quant_conv.weight"].detach().to(dev, f32), sd["post_quant_conv.bias"].detach().to(dev, f32)
        wi, bi = sd[d + "conv_in.conv.weight"].detach().to(dev, f32), sd[d + "conv_in.conv.bias"].detach().to(dev, f32)
        zc = wq.shape[0]
        self.conv_in = _RepConv(torch.einsum("omtyx,mi->oityx", wi, wq.reshape(zc, zc)), bi + torch.einsum("omtyx,m->o", wi, bq), dev)
        self.mid1, self.mid2 = res(d + "mid_block.resnets.0."), res(d + "mid_block.resnets.1.")
        a = d + "mid_block.attentions.0."
        self.attn = {"gn": gn(a + "group_norm"),
                     "wqkv": torch.cat([lin(a + n)[0] for n in ("to_q", "to_k", "to_v")], 0).contiguous(),
                     "bqkv": torch.cat([lin(a + n)[1] for n in ("to_q", "to_k", "to_v")], 0).contiguous(), "proj": lin(a + "to_out.0")}
        self.blocks = []
        blocks, _ = synth.hyvae10_layout(self.cfg)
        for i, (rs, up) in enumerate(blocks):
            u = None
            if up is not None:
                if not up[1]:
                    raise NotImplementedError("time-only up-sampling block (not produced by the 884 layout)")
                p = d + f"up_blocks.{i}.upsamplers.0.conv.conv"
                u = _UpConvNearest(sd[p + ".weight"], sd[p + ".bias"], up[0], dev)
            self.blocks.append(([res(d + f"up_blocks.{i}.resnets.{j}.") for j in range(len(rs))], u))
        self.norm_out = gn(d + "conv_norm_out")
        self.conv_out = rc(d + "conv_out.conv")
        self._ready = True
        return torch.nn.modules.module._IncompatibleKeys([], [])

    @staticmethod
    def _res(d, hold):
        """ResnetBlockCausal3D.forward (unet_causal_3d_blocks.py:455-493).  `hold` is a one-element list that is CONSUMED so the
        block input can be freed as soon as the shortcut exists (61 GB at 720p x 129f x 256 channels)."""
        x = hold.pop()
        T, H, W, C = x.shape
        h = norm_act_conv(x, d["n1"], d["c1"])
        sc = x
        if "sc" in d:                                                        # 1x1x1 conv_shortcut == GEMM over pixels
            sc = ops.gemm(x.reshape(-1, C), d["sc"][0], bias=d["sc"][1]).reshape(T, H, W, -1)
        del x
        # (h + shortcut) / output_scale_factor(=1); the conv epilogue reads the residual and stores the sum at the same address
        return norm_act_conv(h, d["n2"], d["c2"], residual=sc, out=sc)

    def _attn(self, x):
        T, H, W, C = x.shape
        a, N = self.attn, H * W
        xn = a["gn"].apply(x, a["gn"].stats(x), False).reshape(T * N, C)
        buf = torch.zeros(T * N + 8, 3 * C, device=x.device, dtype=bf16)
        qkv = ops.gemm(xn, a["wqkv"], out=buf[:T * N], bias=a["bqkv"])
        npad = (T * N + 63) // 64 * 64
        ws = torch.empty(N * npad * 6, device=x.device, dtype=torch.uint8)
        o = torch.empty(T * N, C, device=x.device, dtype=bf16)
        _lib.call("b200_attention_1head", qkv.data_ptr(), o.data_ptr(), ws.data_ptr(), ws.numel(), T, N, C, float(C) ** -0.5,
                  1 if self.cfg.get("mid_block_causal_attn", True) else 2, _s())
        return ops.gemm(o, a["proj"][0], bias=a["proj"][1], residual=x.reshape(T * N, C)).reshape(T, H, W, C)

    @torch.no_grad()
    def forward(self, z):
        """z [B, zc, T, h, w] -> frames fp32 [B, 3, 4(T-1)+1, 8h, 8w]."""
        if not self._ready:
            raise RuntimeError("HYVAE10Decoder: load_state_dict() must be called before decode")
        outs = []
        zc = self.cfg["latent_channels"]
        for zi in z:
            zi = zi.to(self.device, f32).contiguous()
            _, T, H, W = zi.shape
            zcl = torch.empty(T, H, W, zc, device=self.device, dtype=bf16)
            _lib.call("b200_planar_to_cl", zi.data_ptr(), zcl.data_ptr(), zc, T * H * W, 1, _s())
            hold = [self.conv_in(zcl)]
            hold.append(self._res(self.mid1, hold))
            hold.append(self._attn(hold.pop()))
            hold.append(self._res(self.mid2, hold))
            for rs, up in self.blocks:
                for d in rs:
                    hold.append(self._res(d, hold))
                if up is not None:
                    hold.append(up(hold.pop()))
            outs.append(norm_act_conv(hold.pop(), self.norm_out, self.conv_out, out_mode=2))
        return torch.stack(outs, 0)


HYVAE10Encoder._res = staticmethod(HYVAE10Decoder._res)
HYVAE10Encoder._attn = HYVAE10Decoder._attn


class AutoencoderKLCausal3D(_TiledDecode, torch.nn.Module):
    """Decode surface of models/hyvideo/vae/autoencoder_kl_causal_3d.py::AutoencoderKLCausal3D."""

    def __init__(self, in_channels=3, out_channels=3, down_block_types=None, up_block_types=None, block_out_channels=(128, 256, 512, 512),
                 layers_per_block=2, act_fn="silu", latent_channels=16, norm_num_groups=32, sample_size=256, sample_tsize=64,
                 scaling_factor=0.476986, spatial_compression_ratio=8, time_compression_ratio=4, mid_block_add_attention=True,
                 mid_block_causal_attn=True, device="cuda", **unused):
        super().__init__()
        self.time_compression_ratio, self.spatial_compression_ratio = time_compression_ratio, spatial_compression_ratio
        self.config = types.SimpleNamespace(scaling_factor=scaling_factor, latent_channels=latent_channels,
                                            block_out_channels=tuple(block_out_channels))
        cfg = dict(latent_channels=latent_channels, out_channels=out_channels, block_out_channels=list(block_out_channels),
                   layers_per_block=layers_per_block, norm_num_groups=norm_num_groups, time_compression_ratio=time_compression_ratio,
                   spatial_compression_ratio=spatial_compression_ratio, mid_block_causal_attn=mid_block_causal_attn)
        self.decoder = HYVAE10Decoder(cfg, device)
        self.encoder = HYVAE10Encoder(cfg, device)
        ss = sample_size[0] if isinstance(sample_size, (list, tuple)) else sample_size                    # autoencoder_kl_causal_3d.py:251-262
        self._init_tiles(ss, sample_tsize, int(ss / (2 ** (len(block_out_channels) - 1))), sample_tsize // time_compression_ratio)

    def load_state_dict(self, sd, strict=True, assign=False):
        if "encoder.conv_in.conv.weight" in sd:                 # decode-only checkpoints carry no encoder
            self.encoder.load_state_dict({k: v for k, v in sd.items() if k.startswith(("encoder.", "quant_conv."))})
        return self.decoder.load_state_dict({k: v for k, v in sd.items() if k.startswith(("decoder.", "post_quant_conv."))})

    def decode(self, z, return_dict=True, generator=None):
        out = self._decode_batch(z)
        return types.SimpleNamespace(sample=out) if return_dict else (out,)

    def encode(self, x, return_dict=True):
        """AutoencoderKLCausal3D.encode (autoencoder_kl_causal_3d.py:435-472), tiling off: posterior over quant_conv(encoder(x))."""
        post = _Posterior(self._encode_batch(x))
        return types.SimpleNamespace(latent_dist=post) if return_dict else (post,)
