"""B200-native Wan2.1 VAE decode with the reference `models/wan/modules/vae.py::WanVAE` API
(SURVEY.md section 8b level 3): .decode(zs, tile_size), .decode_to_cpu_uint8(zs, tile_size, ...), .mean/.std/.scale,
.model.z_dim, get_VAE_tile_size.

The reference decodes one latent frame per Python-loop iteration through per-conv feature caches
(vae.py:639-655).  Here the WHOLE clip is decoded in one pass (180 GB of HBM holds every level of a
720p x 81-frame decode), using the exact whole-sequence equivalent of the chunked semantics
(restated and pinned in oracle/vae_oracle.py):
  * CausalConv3d  = causal conv with two leading zero frames  -> TMA out-of-bounds zero fill, no padded copies
  * upsample3d    = frame 0 bypasses time_conv; time_conv runs causally over frames 1.. with zero history and
                    its 2C channels are interleaved in time by the GEMM epilogue's store mapping
Activations are channels-last bf16 [T,H,W,C]; every conv is a tcgen05 implicit GEMM (csrc/gemm_sm100.cuh).
"""
import os

import torch

from .. import _lib, ops, synth

bf16, f32 = torch.bfloat16, torch.float32


def _s():
    return torch.cuda.current_stream().cuda_stream


# the consumer's RMS_norm + SiLU in the producing conv's epilogue where a pixel's channels sit in one CTA (B200_VAE_FUSE_NORM=0: the
# separate norm pass everywhere, for A/B measurements)
FUSE_NORM = os.environ.get("B200_VAE_FUSE_NORM", "1") != "0"
# decoder head conv with its 9 spatial taps stacked into N (B200_VAE_HEAD_STACK=0: the N=16 implicit-GEMM head, for A/B measurements)
HEAD_STACK = os.environ.get("B200_VAE_HEAD_STACK", "1") != "0"


class _Conv:
    """Repacked conv weights: bf16 [Cout, taps, Cin] (tap order t,h,w), fp32 bias."""

    def __init__(self, w, b, device):
        if w.dim() == 4:                                  # Conv2d [Co,Ci,kh,kw] -> kt = 1
            w = w.unsqueeze(2)
        co, ci, kt, kh, kw = w.shape
        self.cout, self.cin, self.k = co, ci, (kt, kh, kw)
        self.w = w.detach().to(device, f32).permute(0, 2, 3, 4, 1).reshape(co, kt * kh * kw, ci).to(bf16).contiguous()
        bias = b.detach().to(device, f32)
        if co % 4:                                        # planar head (Cout=3): pad for safety of vector loads
            bias = torch.cat([bias, bias.new_zeros(16 - co)])
        self.b = bias.contiguous()
        self.w_stack = None
        if HEAD_STACK and co <= 3 and (kt, kh, kw) == (3, 3, 3):
            # decoder head: the 9 spatial taps stacked into the GEMM's N (csrc/vae_ops.cu::b200_conv3d_head_cl):
            # w_stack[(dh*3+dw)*co + c, dt, ci] = w[c, ci, dt, dh, dw], rows padded to 32
            ws = w.detach().to(device, f32).permute(3, 4, 0, 2, 1).reshape(9 * co, kt, ci)
            self.w_stack = torch.cat([ws, ws.new_zeros(32 - 9 * co, kt, ci)], 0).to(bf16).contiguous()

    def head(self, x, T, H, W, prepadded, out=None):
        """Planar fp32 [Cout,T,H,W] = conv(x) through the tap-stacked GEMM + gather; x [T,H,W,C] or the replicate-padded [T+2,H+2,W+2,C]."""
        Hg, Wg = (H + 2, W + 2) if prepadded else (H, W)
        ws = torch.empty(T * Hg * Wg * 32, device=x.device, dtype=f32)
        if out is None:
            out = torch.empty(self.cout, T, H, W, device=x.device, dtype=f32)
        _lib.call("b200_conv3d_head_cl", x.data_ptr(), self.w_stack.data_ptr(), self.b.data_ptr(), ws.data_ptr(), ws.numel() * 4, out.data_ptr(),
                  T, H, W, self.cin, self.cout, int(prepadded), _s())
        return out

    def __call__(self, x, residual=None, out=None, out_mode=0, t_off=0):
        T, H, W, C = x.shape
        assert C == self.cin and x.is_contiguous() and x.dtype == bf16
        kt, kh, kw = self.k
        if out_mode == 2 and self.w_stack is not None and residual is None:
            return self.head(x, T, H, W, False, out)
        if out is None:
            out = (torch.empty(self.cout, T, H, W, device=x.device, dtype=f32) if out_mode == 2
                   else torch.empty(T, H, W, self.cout, device=x.device, dtype=bf16))
        _lib.call("b200_conv3d_cl", x.data_ptr(), self.w.data_ptr(), self.b.data_ptr(),
                  0 if residual is None else residual.data_ptr(), out.data_ptr(), T, H, W, self.cin, self.cout,
                  kt, kh, kw, out_mode, t_off, _s())
        return out


    # ---- fused "this conv -> the consumer's RMS_norm + SiLU" (csrc/conv_sm100.cuh NORM epilogue)
    def fusable(self, W):
        kt, kh, kw = self.k
        return FUSE_NORM and _lib.query("b200_conv_norm_fusable", int(W), self.cin, self.cout, kh, kw) == 1

    def with_norm(self, x, gamma, residual=None, raw=True):
        """-> (conv(x) [+ residual] or None, silu(rms_norm(that) * gamma)): both bf16 [T,H,W,Cout]; raw=False skips the un-normalised
        tensor (a ResidualBlock's first conv: its output is only ever read through the norm, vae.py:246-250)."""
        T, H, W, C = x.shape
        assert C == self.cin and x.is_contiguous() and x.dtype == bf16
        kt, kh, kw = self.k
        out = torch.empty(T, H, W, self.cout, device=x.device, dtype=bf16) if raw else None
        nrm = torch.empty(T, H, W, self.cout, device=x.device, dtype=bf16)
        _lib.call("b200_conv3d_cl_norm", x.data_ptr(), self.w.data_ptr(), self.b.data_ptr(), 0 if residual is None else residual.data_ptr(),
                  0 if out is None else out.data_ptr(), nrm.data_ptr(), gamma.data_ptr(), T, H, W, self.cin, self.cout, kt, kh, kw, _s())
        return out, nrm


class _UpConv:
    """nearest-exact 2x + Conv2d 3x3 (vae.py:124-133) folded into four 2x2 sub-pixel convs (csrc/vae_ops.cu::b200_upconv2x_cl)."""

    def __init__(self, w, b, device):
        co, ci = w.shape[:2]
        w = w.detach().to(device, f32)                      # [Co, Ci, 3, 3]
        self.cout, self.cin = co, ci
        # parity 0 reads source offsets (-1, 0): taps {0} | {1,2};  parity 1 reads (0, +1): taps {0,1} | {2}
        groups = {0: ([0], [1, 2]), 1: ([0, 1], [2])}
        phases = []
        for py in (0, 1):
            for px in (0, 1):
                taps = []
                for a in range(2):
                    for bb in range(2):
                        taps.append(sum(w[:, :, dy, dx] for dy in groups[py][a] for dx in groups[px][bb]))   # [Co, Ci]
                phases.append(torch.stack(taps, 1))           # [Co, 4, Ci]
        self.w4 = torch.stack(phases, 0).to(bf16).contiguous()  # [4, Co, 4, Ci]
        self.b = b.detach().to(device, f32).contiguous()

    def __call__(self, x):
        T, H, W, C = x.shape
        assert C == self.cin and x.is_contiguous() and x.dtype == bf16
        out = torch.empty(T, 2 * H, 2 * W, self.cout, device=x.device, dtype=bf16)
        _lib.call("b200_upconv2x_cl", x.data_ptr(), self.w4.data_ptr(), self.b.data_ptr(), out.data_ptr(), T, H, W, self.cin,
                  self.cout, _s())
        return out


    def fusable(self, W):
        return FUSE_NORM and _lib.query("b200_conv_norm_fusable", int(W), self.cin, self.cout, 2, 2) == 1

    def with_norm(self, x, gamma):
        T, H, W, C = x.shape
        assert C == self.cin and x.is_contiguous() and x.dtype == bf16
        out = torch.empty(T, 2 * H, 2 * W, self.cout, device=x.device, dtype=bf16)
        nrm = torch.empty_like(out)
        _lib.call("b200_upconv2x_cl_norm", x.data_ptr(), self.w4.data_ptr(), self.b.data_ptr(), out.data_ptr(), nrm.data_ptr(), gamma.data_ptr(),
                  T, H, W, self.cin, self.cout, _s())
        return out, nrm


def rms_silu(x, gamma, silu=True, out=None):
    y = torch.empty_like(x) if out is None else out
    assert y.shape == x.shape and y.is_contiguous()
    C = x.shape[-1]
    _lib.call("b200_rms_silu_cl", x.data_ptr(), gamma.data_ptr(), y.data_ptr(), x.numel() // C, C, int(silu), _s())
    return y


def upsample2x(x):
    T, H, W, C = x.shape
    y = torch.empty(T, 2 * H, 2 * W, C, device=x.device, dtype=bf16)
    _lib.call("b200_upsample2x_cl", x.data_ptr(), y.data_ptr(), T, H, W, C, _s())
    return y


class _DownConv:
    """Resample down-sampling (vae.py:134-143): ZeroPad2d((0,1,0,1)) + Conv2d(C, C, 3, stride 2), computed as a stride-1 2x2 conv
    over the space-to-depth tensor [T,H/2,W/2,4C]: kernel row a = 2*da + p (da = s2d tap, p = sub-pixel row), a = 3 gets zero
    weight; the pad row/column on the bottom/right is the TMA zero fill beyond the tensor."""

    def __init__(self, w, b, device, dtype=bf16):
        co, ci = w.shape[:2]
        w = w.detach().to(device, f32)                                      # [Co, Ci, 3, 3]
        w2 = torch.zeros(co, 2, 2, 2, 2, ci, device=device, dtype=f32)      # [Co, da, db, p, q, Ci]
        for da in range(2):
            for p in range(2):
                for db in range(2):
                    for q in range(2):
                        if 2 * da + p < 3 and 2 * db + q < 3:
                            w2[:, da, db, p, q] = w[:, :, 2 * da + p, 2 * db + q]
        self.cout, self.cin = co, ci
        self.w = w2.reshape(co, 4, 4 * ci).to(dtype).contiguous()          # [Co][tap (da,db)][(p,q,c)]
        self.b = b.detach().to(device, f32).contiguous()

    def __call__(self, x, spare_frame=False):
        """x [T,H,W,C] -> [T(+1),H/2,W/2,Co]; spare_frame allocates one extra (unwritten) frame so the result can be viewed as
        frame pairs by _TimeDownConv."""
        T, H, W, C = x.shape
        if H % 2 or W % 2:
            raise ValueError("VAE encode: frame height and width must be even at every level (multiples of 8)")
        s2d = torch.empty(T, H // 2, W // 2, 4 * C, device=x.device, dtype=bf16)
        _lib.call("b200_space_to_depth_cl", x.data_ptr(), s2d.data_ptr(), T, H, W, C, _s())
        out = torch.empty(T + int(spare_frame), H // 2, W // 2, self.cout, device=x.device, dtype=bf16)
        h, w = H // 2, W // 2
        _lib.call("b200_conv3d_cl_view", s2d.data_ptr(), T, h, w, 0, 0, 0, self.w.data_ptr(), self.b.data_ptr(), 0, out.data_ptr(),
                  T, h, w, 4 * C, self.cout, 1, 2, 2, h * w * self.cout, w * self.cout, self.cout, _s())
        return out


class _TimeDownConv:
    """'downsample3d' time_conv (vae.py:141-143, 190-212): CausalConv3d(C, C, (3,1,1), stride (2,1,1), no padding) run chunk by
    chunk with a one-frame cache.  Over the whole sequence: out[0] = y[0]; out[j] = w0 y[2j-2] + w1 y[2j-1] + w2 y[2j].  With y
    viewed as frame PAIRS stacked along H ([m+1, 2h, w, C]) the even frames and the odd frames are two windows of one tensor:
    launch 1 = 2-tap conv (w0, w2) over the even window, launch 2 = 1-tap conv (w1) over the odd window accumulated through the
    residual input."""

    def __init__(self, w, b, device, dtype=bf16):
        w = w.detach().to(device, f32)[:, :, :, 0, 0]                        # [Co, Ci, 3]
        self.cout, self.cin = w.shape[:2]
        self.w_even = torch.stack([w[:, :, 0], w[:, :, 2]], 1).to(dtype).contiguous()      # [Co, 2, Ci]
        self.w_odd = w[:, :, 1].unsqueeze(1).to(dtype).contiguous()                        # [Co, 1, Ci]
        self.b = b.detach().to(device, f32).contiguous()

    def __call__(self, y, T):
        """y: buffer of T+1 frames [T+1,h,w,C] whose first T frames are valid (T odd)."""
        _, h, w, C = y.shape
        m = (T - 1) // 2
        out = torch.empty(1 + m, h, w, self.cout, device=y.device, dtype=bf16)
        out[0].copy_(y[0])                                                   # the first chunk's frame bypasses time_conv (:196-198)
        if m > 0:
            o1 = out[1:]
            st = (h * w * self.cout, w * self.cout, self.cout)
            _lib.call("b200_conv3d_cl_view", y.data_ptr(), m + 1, 2 * h, w, 0, 0, 0, self.w_even.data_ptr(), self.b.data_ptr(), 0,
                      o1.data_ptr(), m, h, w, C, self.cout, 2, 1, 1, *st, _s())
            _lib.call("b200_conv3d_cl_view", y.data_ptr(), m + 1, 2 * h, w, 0, h, 0, self.w_odd.data_ptr(), 0, o1.data_ptr(),
                      o1.data_ptr(), m, h, w, C, self.cout, 1, 1, 1, *st, _s())
        return out


class WanVAEDecoder(torch.nn.Module):
    """Decoder half of WanVAE_ (vae.py:549-662); `decode(z, scale)` mirrors WanVAE_.decode."""

    def __init__(self, cfg=None, device="cuda"):
        super().__init__()
        self.cfg = dict(cfg or synth.VAE_CFG)
        self.z_dim = self.cfg["z_dim"]
        self.upsampler_factor = 1
        self.device = torch.device(device)
        self._ready = False

    def load_state_dict(self, sd, strict=True, assign=False):
        dev = self.device
        g = lambda k: sd[k].detach().to(dev, f32).contiguous()  # noqa: E731
        self.conv2_w, self.conv2_b = g("conv2.weight").reshape(self.z_dim, self.z_dim), g("conv2.bias")
        self.conv1 = _Conv(sd["decoder.conv1.weight"], sd["decoder.conv1.bias"], dev)

        def res(p):
            d = {"g0": g(p + "residual.0.gamma").reshape(-1), "c0": _Conv(sd[p + "residual.2.weight"], sd[p + "residual.2.bias"], dev),
                 "g1": g(p + "residual.3.gamma").reshape(-1), "c1": _Conv(sd[p + "residual.6.weight"], sd[p + "residual.6.bias"], dev)}
            if p + "shortcut.weight" in sd:
                d["sc"] = _Conv(sd[p + "shortcut.weight"], sd[p + "shortcut.bias"], dev)
            return d

        self.mid0, self.mid2 = res("decoder.middle.0."), res("decoder.middle.2.")
        a = "decoder.middle.1."
        c0 = sd[a + "proj.weight"].shape[0]
        self.attn = {"g": g(a + "norm.gamma").reshape(-1), "wqkv": sd[a + "to_qkv.weight"].detach().to(dev, bf16).reshape(3 * c0, c0).contiguous(),
                     "bqkv": g(a + "to_qkv.bias"), "wproj": sd[a + "proj.weight"].detach().to(dev, bf16).reshape(c0, c0).contiguous(),
                     "bproj": g(a + "proj.bias")}
        _, ups, _ = synth.vae_decoder_layout(self.cfg)
        self.ups = []
        for j, u in enumerate(ups):
            p = f"decoder.upsamples.{j}."
            if u[0] == "res":
                self.ups.append(("res", res(p)))
            else:
                d = {"conv": _UpConv(sd[p + "resample.1.weight"], sd[p + "resample.1.bias"], dev)}
                if u[0] == "up3d":
                    d["time"] = _Conv(sd[p + "time_conv.weight"], sd[p + "time_conv.bias"], dev)
                self.ups.append((u[0], d))
        self.head_g = g("decoder.head.0.gamma").reshape(-1)
        self.head = _Conv(sd["decoder.head.2.weight"], sd["decoder.head.2.bias"], dev)
        self._ready = True
        return torch.nn.modules.module._IncompatibleKeys([], [])

    # ---- blocks
    @staticmethod
    def _res(d, x, xn=None, next_gamma=None):
        """ResidualBlock (vae.py:238-273) -> (x', silu(norm(x') * next_gamma) or None).  xn: silu(norm(x) * g0) if the producer of x
        already wrote it; next_gamma: norm weight of the layer that consumes x' (the next block's residual.0, or the head's norm)."""
        h = d["sc"](x) if "sc" in d else x                                   # vae.py:255 shortcut
        if xn is None:
            xn = rms_silu(x, d["g0"])
        W = x.shape[2]
        if d["c0"].fusable(W):
            yn = d["c0"].with_norm(xn, d["g1"], raw=False)[1]                # the raw conv output is only read through the norm
        else:
            yn = rms_silu(d["c0"](xn), d["g1"])
        if next_gamma is not None and d["c1"].fusable(W):
            return d["c1"].with_norm(yn, next_gamma, residual=h)             # vae.py:273 x + h, plus the consumer's norm
        return d["c1"](yn, residual=h), None

    def _attn(self, x):
        T, H, W, C = x.shape
        a = self.attn
        xn = rms_silu(x, a["g"], silu=False).reshape(T * H * W, C)
        N = H * W
        buf = torch.zeros(T * N + 8, 3 * C, device=x.device, dtype=bf16)    # 8 spare rows: key tail when N % 8 != 0
        qkv = ops.gemm(xn, a["wqkv"], out=buf[:T * N], bias=a["bqkv"])      # 1x1 conv == GEMM over pixels
        npad = (N + 63) // 64 * 64
        ws = torch.empty(N * npad * 6, device=x.device, dtype=torch.uint8)
        o = torch.empty(T * N, C, device=x.device, dtype=bf16)
        _lib.call("b200_attention_1head", qkv.data_ptr(), o.data_ptr(), ws.data_ptr(), ws.numel(), T, N, C, float(C) ** -0.5, 0, _s())
        return ops.gemm(o, a["wproj"], bias=a["bproj"], residual=x.reshape(T * N, C)).reshape(T, H, W, C)

    @staticmethod
    def _up(kind, d, x, next_gamma=None):
        T, H, W, C = x.shape
        if kind == "up3d" and T > 1:
            y = torch.empty(2 * T - 1, H, W, C, device=x.device, dtype=bf16)
            y[0].copy_(x[0])                                                 # frame 0 bypasses time_conv (vae.py:155-158)
            d["time"](x[1:], out=y, out_mode=1, t_off=1)                     # causal over frames 1.., interleaved store
            x = y
        if next_gamma is not None and d["conv"].fusable(W):
            return d["conv"].with_norm(x, next_gamma)
        return d["conv"](x), None                                       # 2x nearest + 3x3 conv as sub-pixel convs

    @torch.no_grad()
    def decode_frames(self, z, mean, std):
        """z [16,Tl,h,w] fp32 on device -> fp32 planar frames [3, 4(Tl-1)+1, 8h, 8w] (un-clamped)."""
        if not self._ready:
            raise RuntimeError("WanVAE: load_state_dict() must be called before decode")
        C, T, H, W = z.shape
        z = z.to(self.device, f32).contiguous()
        x = torch.empty(T, H, W, 16, device=self.device, dtype=bf16)
        _lib.call("b200_vae_prologue", z.data_ptr(), mean.data_ptr(), std.data_ptr(), self.conv2_w.data_ptr(),
                  self.conv2_b.data_ptr(), x.data_ptr(), T, H, W, _s())
        x = self.conv1(x)
        x = self._res(self.mid0, x)[0]
        x = self._attn(x)
        x = self._res(self.mid2, x)[0]
        xn = None                                       # silu(norm(x)) when the producer of x already wrote it
        for idx, (kind, d) in enumerate(self.ups):
            nxt = self.ups[idx + 1] if idx + 1 < len(self.ups) else None
            next_gamma = self.head_g if nxt is None else (nxt[1]["g0"] if nxt[0] == "res" else None)
            x, xn = self._res(d, x, xn, next_gamma) if kind == "res" else self._up(kind, d, x, next_gamma)
        if xn is None:
            xn = rms_silu(x, self.head_g)
        return self.head(xn, out_mode=2)

    # ------------------------------------------------------------------ time-sliced (streaming) decode
    # The whole-clip decode above keeps ~4 full-resolution activations live (0.7 GB per latent frame each at 720p): beyond ~40 latent
    # frames at 720p (or ~16 at 1080p) that exceeds the 180 GB of a B200.  The streamed decode walks the latent frames in slices of
    # `chunk` frames and carries, for every temporal conv, the last kt-1 = 2 frames of its input to the next slice -- the reference's
    # chunked decode with per-conv feature caches (vae.py:639-655, CausalConv3d.forward :55-61), for slices of any length instead of one
    # latent frame.  Every output value sees exactly the operands of the whole-clip decode, so the two are bit-identical.
    class _Hist:
        def __init__(self, dev):
            self.h, self.dev = {}, dev

        def ext(self, key, K, H, W, C):
            """[2 + K, H, W, C] buffer whose first 2 frames are the history of conv `key` (zeros before the first slice)."""
            xe = torch.empty(K + 2, H, W, C, device=self.dev, dtype=bf16)
            prev = self.h.get(key)
            if prev is None:
                xe[:2].zero_()
            else:
                xe[:2].copy_(prev)
            return xe

        def keep(self, key, xe):
            self.h[key] = xe[-2:].clone()

    @staticmethod
    def _sconv(conv, xe, K, residual=None, out=None, nrm=None, gamma=None, out_mode=0, t_off=0):
        _, H, W, C = xe.shape
        kt, kh, kw = conv.k
        assert kt == 3 and xe.shape[0] == K + 2 and C == conv.cin
        _lib.call("b200_conv3d_cl_stream", xe.data_ptr(), conv.w.data_ptr(), conv.b.data_ptr(), 0 if residual is None else residual.data_ptr(),
                  0 if out is None else out.data_ptr(), 0 if nrm is None else nrm.data_ptr(), 0 if gamma is None else gamma.data_ptr(),
                  K, H, W, conv.cin, conv.cout, kt, kh, kw, out_mode, t_off, _s())

    def _res_s(self, st, key, d, x, xe0, nxt):
        """Streaming ResidualBlock: x raw [K,H,W,C]; xe0 = history-extended normalised input if the producer wrote it; nxt = (key, gamma)
        of the consumer's conv or None.  -> (x' raw, history-extended normalised x' or None)."""
        K, H, W, C = x.shape
        h = d["sc"](x) if "sc" in d else x
        if xe0 is None:
            xe0 = st.ext(key + ".c0", K, H, W, C)
            rms_silu(x, d["g0"], out=xe0[2:])
        st.keep(key + ".c0", xe0)
        co = d["c0"].cout
        xe1 = st.ext(key + ".c1", K, H, W, co)
        if d["c0"].fusable(W):
            self._sconv(d["c0"], xe0, K, nrm=xe1[2:], gamma=d["g1"])
        else:
            y = torch.empty(K, H, W, co, device=x.device, dtype=bf16)
            self._sconv(d["c0"], xe0, K, out=y)
            rms_silu(y, d["g1"], out=xe1[2:])
        del xe0
        st.keep(key + ".c1", xe1)
        x2 = torch.empty(K, H, W, co, device=x.device, dtype=bf16)
        xe_n = None
        if nxt is not None and d["c1"].fusable(W):
            xe_n = st.ext(nxt[0], K, H, W, co)
            self._sconv(d["c1"], xe1, K, residual=h, out=x2, nrm=xe_n[2:], gamma=nxt[1])
        else:
            self._sconv(d["c1"], xe1, K, residual=h, out=x2)
        return x2, xe_n

    def _up_s(self, st, key, kind, d, x, first, nxt):
        K, H, W, C = x.shape
        if kind == "up3d":
            if first:
                y = torch.empty(2 * K - 1, H, W, C, device=x.device, dtype=bf16)
                y[0].copy_(x[0])                                             # latent frame 0 bypasses time_conv (vae.py:155-158)
                d["time"](x[1:], out=y, out_mode=1, t_off=1)                 # zero history: frame 0 is never seen (vae.py:179-180)
                st.h[key + ".time"] = x[-2:].clone()
            else:
                xe = st.ext(key + ".time", K, H, W, C)
                xe[2:].copy_(x)
                st.keep(key + ".time", xe)
                y = torch.empty(2 * K, H, W, C, device=x.device, dtype=bf16)
                self._sconv(d["time"], xe, K, out=y, out_mode=1, t_off=0)
            x = y
        T2 = x.shape[0]
        if nxt is not None and d["conv"].fusable(W):
            co = d["conv"].cout
            x2 = torch.empty(T2, 2 * H, 2 * W, co, device=x.device, dtype=bf16)
            xe_n = st.ext(nxt[0], T2, 2 * H, 2 * W, co)
            _lib.call("b200_upconv2x_cl_norm", x.data_ptr(), d["conv"].w4.data_ptr(), d["conv"].b.data_ptr(), x2.data_ptr(), xe_n[2:].data_ptr(),
                      nxt[1].data_ptr(), T2, H, W, d["conv"].cin, co, _s())
            return x2, xe_n
        return d["conv"](x), None

    @torch.no_grad()
    def decode_frames_streamed(self, z, mean, std, chunk=8, out=None):
        """Time-sliced decode_frames: same result (bit-identical), memory bounded by `chunk` latent frames per slice (>= 3)."""
        if not self._ready:
            raise RuntimeError("WanVAE: load_state_dict() must be called before decode")
        C, T, H, W = z.shape
        chunk = max(3, int(chunk))
        if T <= chunk or self.head.w_stack is None:
            return self.decode_frames(z, mean, std)
        z = z.to(self.device, f32).contiguous()
        F = 4 * (T - 1) + 1
        if out is None:
            out = torch.empty(3, F, 8 * H, 8 * W, device=self.device, dtype=f32)
        st = self._Hist(self.device)
        nups = len(self.ups)
        f0, t0 = 0, 0
        while t0 < T:
            K = min(chunk, T - t0)
            if T - (t0 + K) in (1, 2):          # never leave a tail of fewer than 3 latent frames
                K = T - t0
            first = t0 == 0
            zc = z[:, t0:t0 + K].contiguous()
            xe = st.ext("conv1", K, H, W, 16)
            _lib.call("b200_vae_prologue", zc.data_ptr(), mean.data_ptr(), std.data_ptr(), self.conv2_w.data_ptr(),
                      self.conv2_b.data_ptr(), xe[2:].data_ptr(), K, H, W, _s())
            st.keep("conv1", xe)
            x = torch.empty(K, H, W, self.conv1.cout, device=self.device, dtype=bf16)
            self._sconv(self.conv1, xe, K, out=x)
            x = self._res_s(st, "mid0", self.mid0, x, None, None)[0]
            x = self._attn(x)                                                   # per frame: no history
            x = self._res_s(st, "mid2", self.mid2, x, None, None)[0]
            xe_n = None
            for idx, (kind, d) in enumerate(self.ups):
                nk = self.ups[idx + 1][0] if idx + 1 < nups else "head"
                nxt = ("head", self.head_g) if nk == "head" else ((f"up{idx + 1}.c0", self.ups[idx + 1][1]["g0"]) if nk == "res" else None)
                if kind == "res":
                    x, xe_n = self._res_s(st, f"up{idx}", d, x, xe_n, nxt)
                else:
                    x, xe_n = self._up_s(st, f"up{idx}", kind, d, x, first, nxt)
            n, Hf, Wf, Ch = x.shape
            if xe_n is None:
                xe_n = st.ext("head", n, Hf, Wf, Ch)
                rms_silu(x, self.head_g, out=xe_n[2:])
            st.keep("head", xe_n)
            del x
            fr = self.head.head(xe_n, n, Hf, Wf, 2)                            # streaming slice: 2 history frames in front
            out[:, f0:f0 + n].copy_(fr)
            del fr, xe_n
            f0 += n
            t0 += K
        assert f0 == F
        return out

    def decode(self, z, scale=None, any_end_frame=False):
        """WanVAE_.decode contract (vae.py:628-662): z [1,16,T,h,w]; scale = [mean, 1/std] -> [1,3,F,H,W] fp32."""
        if any_end_frame:
            raise NotImplementedError("any_end_frame decode is outside the t2v/i2v2_2 hot path")
        mean, inv_std = scale
        mean = torch.as_tensor(mean, dtype=f32, device=self.device).contiguous()
        std = (1.0 / torch.as_tensor(inv_std, dtype=f32, device=self.device)).contiguous()
        return torch.stack([self.decode_frames(zi, mean, std) for zi in z], 0)


class WanVAEEncoder(torch.nn.Module):
    """Encoder half of WanVAE_ (vae.py:318-427, 586-625), whole-sequence: `encode(x, scale)` mirrors WanVAE_.encode.  The reference
    feeds the encoder chunks of 1, 4, 4, ... frames with per-conv feature caches; that equals causal convs over the whole
    sequence plus the frame-pair rule of _TimeDownConv (pinned at 4e-7 by oracle/vae_oracle.py::vae_encode against the
    reference).  conv1 (1x1x1), the mu half of `chunk(2)` and the latent normalisation are folded into the head conv."""

    def __init__(self, cfg=None, device="cuda"):
        super().__init__()
        self.cfg = dict(cfg or synth.VAE_CFG)
        self.z_dim = self.cfg["z_dim"]
        self.device = torch.device(device)
        self._ready, self._head = False, {}

    def load_state_dict(self, sd, strict=True, assign=False):
        dev = self.device
        g = lambda k: sd[k].detach().to(dev, f32).contiguous()  # noqa: E731
        w1 = sd["encoder.conv1.weight"].detach().to(dev, f32)
        self.conv1 = _Conv(torch.cat([w1, w1.new_zeros(w1.shape[0], 5, *w1.shape[2:])], 1), sd["encoder.conv1.bias"], dev)   # Cin 3 -> 8

        def res(p):
            d = {"g0": g(p + "residual.0.gamma").reshape(-1), "c0": _Conv(sd[p + "residual.2.weight"], sd[p + "residual.2.bias"], dev),
                 "g1": g(p + "residual.3.gamma").reshape(-1), "c1": _Conv(sd[p + "residual.6.weight"], sd[p + "residual.6.bias"], dev)}
            if p + "shortcut.weight" in sd:
                d["sc"] = _Conv(sd[p + "shortcut.weight"], sd[p + "shortcut.bias"], dev)
            return d
        _, downs, _ = synth.vae_encoder_layout(self.cfg)
        self.downs = []
        for j, u in enumerate(downs):
            p = f"encoder.downsamples.{j}."
            if u[0] == "res":
                self.downs.append(("res", res(p)))
            else:
                d = {"conv": _DownConv(sd[p + "resample.1.weight"], sd[p + "resample.1.bias"], dev)}
                if u[0] == "down3d":
                    d["time"] = _TimeDownConv(sd[p + "time_conv.weight"], sd[p + "time_conv.bias"], dev)
                self.downs.append((u[0], d))
        self.mid0, self.mid2 = res("encoder.middle.0."), res("encoder.middle.2.")
        a = "encoder.middle.1."
        c0 = sd[a + "proj.weight"].shape[0]
        self.attn = {"g": g(a + "norm.gamma").reshape(-1), "wqkv": sd[a + "to_qkv.weight"].detach().to(dev, bf16).reshape(3 * c0, c0).contiguous(),
                     "bqkv": g(a + "to_qkv.bias"), "wproj": sd[a + "proj.weight"].detach().to(dev, bf16).reshape(c0, c0).contiguous(),
                     "bproj": g(a + "proj.bias")}
        self.head_g = g("encoder.head.0.gamma").reshape(-1)
        self.head_w, self.head_b = g("encoder.head.2.weight"), g("encoder.head.2.bias")
        self.c1_w, self.c1_b = g("conv1.weight").reshape(2 * self.z_dim, 2 * self.z_dim), g("conv1.bias")
        self._ready, self._head = True, {}
        return torch.nn.modules.module._IncompatibleKeys([], [])

    def _head_conv(self, mean, inv_std):
        """head conv3d -> conv1 1x1x1 -> mu = chunk(2)[0] -> (mu - mean) * inv_std (vae.py:612-619) as ONE conv with fp32 planar output:
        W = diag(inv_std) W1[:z] Wh,  b = inv_std * (W1[:z] bh + b1[:z] - mean)."""
        key = (tuple(mean.tolist()), tuple(inv_std.tolist()))
        if key not in self._head:
            z = self.z_dim
            m = inv_std[:, None] * self.c1_w[:z]                              # [z, 2z]
            w = torch.einsum("om,mctyx->octyx", m, self.head_w)
            b = inv_std * (self.c1_w[:z] @ self.head_b + self.c1_b[:z] - mean)
            self._head = {key: _Conv(w, b, self.device)}
        return self._head[key]

    _res = staticmethod(WanVAEDecoder._res)
    _attn = WanVAEDecoder._attn

    @torch.no_grad()
    def encode_frames(self, x, mean, inv_std):
        """x [3, 1+4k, H, W] fp32 in [-1,1] on device -> normalised latent mean [16, 1+k, H/8, W/8] fp32."""
        if not self._ready:
            raise RuntimeError("WanVAE: load_state_dict() must be called before encode")
        C, T, H, W = x.shape
        if (T - 1) % 4:
            raise ValueError("VAE encode takes 1 + 4k frames (vae.py:590-594)")
        if H % 8 or W % 8:
            raise ValueError("VAE encode: frame height and width must be multiples of 8")
        x = x.to(self.device, f32).contiguous()
        h = torch.empty(T, H, W, 8, device=self.device, dtype=bf16)
        _lib.call("b200_planar_to_cl_pad", x.data_ptr(), h.data_ptr(), C, T * H * W, 8, _s())
        h = self.conv1(h)
        for kind, d in self.downs:
            if kind == "res":
                h = self._res(d, h)[0]
            else:
                t = h.shape[0]
                temporal = kind == "down3d" and t > 1
                h = d["conv"](h, spare_frame=temporal)
                if temporal:
                    h = d["time"](h, t)
        h = self._res(self.mid0, h)[0]
        h = self._attn(h)
        h = self._res(self.mid2, h)[0]
        return self._head_conv(mean, inv_std)(rms_silu(h, self.head_g), out_mode=2)

    def encode(self, x, scale=None, any_end_frame=False):
        """WanVAE_.encode contract (vae.py:586-625): x [1,3,T,H,W]; scale = [mean, 1/std] -> mu [1,16,1+(T-1)/4,H/8,W/8] fp32."""
        if any_end_frame:
            raise NotImplementedError("any_end_frame encode is outside the t2v/i2v2_2 hot path")
        z = self.z_dim
        mean = torch.zeros(z) if scale is None else torch.as_tensor(scale[0], dtype=f32).reshape(-1).cpu()
        inv_std = torch.ones(z) if scale is None else torch.as_tensor(scale[1], dtype=f32).reshape(-1).cpu()
        mean, inv_std = mean.expand(z).to(self.device, f32).contiguous(), inv_std.expand(z).to(self.device, f32).contiguous()
        return torch.stack([self.encode_frames(xi, mean, inv_std) for xi in x], 0)


def _blend(a, b, extent, vertical):
    """blend_v / blend_h (vae.py:664-674) in place on tile b; a, b fp32 planar [C,F,h,w]."""
    _lib.call("b200_blend_edge_f32", a.data_ptr(), b.data_ptr(), a.shape[0] * a.shape[1], a.shape[2], a.shape[3], b.shape[2], b.shape[3],
              int(extent), int(vertical), _s())
    return b


def spatial_tiles(x, tile, stride, fn, blend_extent, row_limit):
    """The tiling loop shared by spatial_tiled_decode (vae.py:676-723) and spatial_tiled_encode (:841-881): `fn` maps the crop
    x[..., i:i+tile, j:j+tile] to a contiguous fp32 [C,F,h',w'] tile; each tile is cross-faded with its (already blended) upper
    and left neighbours over `blend_extent` rows / columns, cropped to `row_limit` and written into the result."""
    H, W = x.shape[-2:]
    rows = [[fn(x[..., i:i + tile, j:j + tile].contiguous()) for j in range(0, W, stride)] for i in range(0, H, stride)]
    out_rows = []
    for i, row in enumerate(rows):
        out = []
        for j, t in enumerate(row):
            if i > 0:
                _blend(rows[i - 1][j], t, blend_extent, True)
            if j > 0:
                _blend(row[j - 1], t, blend_extent, False)
            out.append(t[:, :, :row_limit, :row_limit])
        out_rows.append(torch.cat(out, -1))
    return torch.cat(out_rows, -2)


class WanVAE:
    """Mirror of models/wan/modules/vae.py::WanVAE (:935-1027)."""

    def __init__(self, z_dim=16, vae_pth=None, dtype=torch.float, upsampler_factor=1, device="cuda", preprocess_sd=None,
                 state_dict=None, cfg=None):
        if upsampler_factor != 1:
            raise NotImplementedError("upsampler VAE variants are outside the hot path")
        self.dtype, self.device, self.z_dim = dtype, device, z_dim
        self.mean = torch.tensor(synth.VAE_MEAN, dtype=f32, device=device)          # vae.py:948-951
        self.std = torch.tensor(synth.VAE_STD, dtype=f32, device=device)            # vae.py:952-955
        self.scale = [self.mean, 1.0 / self.std]
        self.model = WanVAEDecoder(cfg, device)
        self.encoder_model = WanVAEEncoder(cfg, device)
        if state_dict is None and vae_pth is not None:
            state_dict = torch.load(vae_pth, map_location="cpu")
            if preprocess_sd is not None:
                state_dict = preprocess_sd(state_dict)
        if state_dict is not None:
            self.model.load_state_dict(state_dict)
            if "encoder.conv1.weight" in state_dict:           # decode-only checkpoints / synthetic decoders carry no encoder
                self.encoder_model.load_state_dict(state_dict)

    @staticmethod
    def get_VAE_tile_size(vae_config, device_mem_capacity, mixed_precision, output_height=None, output_width=None):
        """vae.py:968-1000 picks a tile size from free VRAM; on a 180 GB B200 the whole clip is decoded untiled."""
        return 0

    def _tiled_decode(self, u, tile_size):
        """WanVAE_.spatial_tiled_decode (vae.py:676-723): latent tiles of tile_size/8 with 25 % overlap, each decoded as a whole clip,
        seams cross-faded.  get_VAE_tile_size() returns 0 on a B200 (whole clip un-tiled); an explicit tile_size > 0 (outputs beyond
        1080p, vae.py:975-978) reproduces the reference's tiled result."""
        tile_size = int(tile_size)
        if tile_size < 16:
            raise ValueError("VAE tile_size must be >= 16 pixels")
        tl = int(tile_size / 8)
        mean, std = self.mean, self.std
        fn = lambda z: self.model.decode_frames(z, mean, std)                                        # noqa: E731
        return spatial_tiles(u.to(self.device, f32), tl, int(tl * 0.75), fn, int(tile_size * 0.25), tile_size - int(tile_size * 0.25))

    def decode(self, zs, tile_size=0, any_end_frame=False):
        """list of [16,Tl,h,w] -> list of fp32 [3,F,H,W] clamped to [-1,1] (vae.py:1012-1017)."""
        if tile_size and tile_size > 0:
            if any_end_frame:
                raise NotImplementedError("any_end_frame decode is outside the t2v/i2v2_2 hot path")
            return [self._tiled_decode(u, tile_size).clamp_(-1, 1) for u in zs]
        return [self.model.decode(u.unsqueeze(0), self.scale, any_end_frame)[0].clamp_(-1, 1) for u in zs]

    def decode_to_cpu_uint8(self, zs, tile_size=0, target_frames=None, target_height=None, target_width=None,
                            any_end_frame=False, frame_start=0):
        """vae.py:1021-1027 -> :741-767: uint8 CPU frames, round(clamp((clamp(x,-1,1)+1)*127.5, 0, 255))."""
        outs = []
        for u in zs:
            # tile_size > 0: same tiles / seams as the reference's streaming tiled writer (vae.py:741-839), blended in fp32 first
            fr = self._tiled_decode(u, tile_size) if tile_size and tile_size > 0 else self.model.decode(u.unsqueeze(0), self.scale, any_end_frame)[0]
            n = fr.shape[1]
            fs = min(max(0, int(frame_start or 0)), n)
            fe = n if target_frames is None else min(n, fs + int(target_frames))
            # quantise the whole frame range (H = 8h, W = 8w: numel % 4 == 0 as the 16-byte kernel needs), crop the uint8 result
            # afterwards: the reference accepts any target_height / target_width (vae.py:757-767)
            fr = fr[:, fs:fe].contiguous()
            u8 = torch.empty(fr.shape, device=fr.device, dtype=torch.uint8)
            if fr.numel():
                _lib.call("b200_frames_to_u8", fr.data_ptr(), u8.data_ptr(), fr.numel(), _s())
            outs.append(u8[:, :, :target_height, :target_width].cpu())
        return outs

    def encode(self, videos, tile_size=0, any_end_frame=False):
        """list of [3,T,H,W] videos in [-1,1] -> list of normalised latents fp32 [16,1+(T-1)/4,H/8,W/8] (vae.py:1002-1010), un-tiled:
        the reference's default tile_size=256 selects spatial_tiled_encode (tile + blend to fit small GPUs); the default here is the
        whole clip at once = the reference's tile_size=0 branch; tile_size > 0 reproduces the tiled branch."""
        if tile_size and tile_size > 0:
            # WanVAE_.spatial_tiled_encode (vae.py:841-881): video tiles of tile_size with 25 % overlap, latents cross-faded.  The
            # latent normalisation is affine and the cross-fade weights sum to 1, so normalising each tile first is the same map.
            if any_end_frame:
                raise NotImplementedError("any_end_frame encode is outside the t2v/i2v2_2 hot path")
            tile_size = int(tile_size)
            if tile_size < 32 or tile_size % 8:
                raise ValueError("VAE encode tile_size must be a multiple of 8, >= 32")
            tl = tile_size // 8
            mean, inv_std = self.mean.to(self.device), (1.0 / self.std).to(self.device)
            fn = lambda x: self.encoder_model.encode_frames(x, mean, inv_std)                        # noqa: E731
            return [spatial_tiles(u.to(self.device, f32), tile_size, int(tile_size * 0.75), fn, int(tl * 0.25), tl - int(tl * 0.25))
                    for u in videos]
        return [self.encoder_model.encode(u.unsqueeze(0), self.scale, any_end_frame)[0] for u in videos]
