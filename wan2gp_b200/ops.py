"""Thin torch-tensor wrappers over the C ABI.  PyTorch is only the allocator / stream provider here:
every function checks dtype/layout, passes raw device pointers + the current CUDA stream, and raises
B200Error on failure (no eager fallback)."""
import math

import ctypes

import torch

from . import _lib

bf16, f32 = torch.bfloat16, torch.float32
CFG_DOTS_FLOATS = 2 + 2 * 1184 + 2        # include/wan2gp_b200.h::B200_CFG_DOTS_FLOATS


def _stream():
    return torch.cuda.current_stream().cuda_stream


# Optional per-kernel timing with CUDA events on the launching stream (bench.py's roofline leg).  Disabled by default:
# when `TIMED` holds a kernel name, every launch of it is bracketed by two event records (no synchronisation here).
TIMED = {}          # name -> list of (start_event, end_event, work) ; filled only for names present as keys


def _timed(name, work, fn):
    rec = TIMED.get(name)
    if rec is None:
        return fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = fn()
    e1.record()
    rec.append((e0, e1, work))
    return out


def _p(t):
    return 0 if t is None else t.data_ptr()


def _chk(t, dtype, name):
    if t.device.type != "cuda":
        raise _lib.B200Error(f"{name}: expected a CUDA tensor (there is no CPU path)")
    if t.dtype != dtype:
        raise _lib.B200Error(f"{name}: expected {dtype}, got {t.dtype}")


def gemm(a, b, out=None, bias=None, gate=None, residual=None, act=0, out_dtype=bf16, accumulate=False, b_mn_major=False):
    """out[M,N] = act(a[M,K] @ b[N,K]^T + bias) * gate   (b_mn_major: b is [K,N])."""
    _chk(a, bf16, "a"), _chk(b, bf16, "b")
    assert a.dim() == 2 and b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1
    M, K = a.shape
    N = b.shape[1] if b_mn_major else b.shape[0]
    assert (b.shape[0] if b_mn_major else b.shape[1]) == K
    if out is None:
        out = torch.empty(M, N, device=a.device, dtype=out_dtype)
    assert out.shape == (M, N) and out.stride(1) == 1 and out.dtype in (bf16, f32)
    for v, n in ((bias, "bias"), (gate, "gate")):
        if v is not None:
            _chk(v, f32, n)
            assert v.numel() == N and v.is_contiguous()
    if residual is not None:
        _chk(residual, bf16, "residual")
        assert residual.shape == (M, N) and residual.stride() == out.stride() and out.dtype == bf16
    _timed("gemm", 2.0 * M * N * K, lambda: _lib.call(
        "b200_gemm_bf16", a.data_ptr(), b.data_ptr(), out.data_ptr(), M, N, K, a.stride(0), b.stride(0), out.stride(0),
        _p(bias), _p(gate), _p(residual), int(act), int(out.dtype == f32), int(accumulate), int(b_mn_major), _stream()))
    return out


def ln_modulate(x, shift, scale, affine=False, eps=1e-6, out=None, pre_round=False):
    _chk(x, f32, "x"), _chk(shift, f32, "shift"), _chk(scale, f32, "scale")
    L, D = x.shape
    assert x.is_contiguous() and shift.numel() == D and scale.numel() == D
    if out is None:
        out = torch.empty(L, D, device=x.device, dtype=bf16)
    _lib.call("b200_ln_modulate", x.data_ptr(), shift.data_ptr(), scale.data_ptr(), int(affine), int(pre_round), out.data_ptr(),
              L, D, float(eps), _stream())
    return out


def rmsnorm_rope_(x, w, eps=1e-6, cos=None, sin=None, per_head=False):
    """In place on bf16 x[L, D] (last dim contiguous, arbitrary row stride); per_head: norm over each 128-wide head, w [128]."""
    _chk(x, bf16, "x"), _chk(w, f32, "w")
    L, D = x.shape
    assert x.stride(1) == 1 and w.numel() == (128 if per_head else D)
    if cos is not None:
        _chk(cos, f32, "cos"), _chk(sin, f32, "sin")
        assert cos.shape == (L, 128) and sin.shape == (L, 128) and cos.is_contiguous() and sin.is_contiguous()
    _lib.call("b200_rmsnorm_rope", x.data_ptr(), x.stride(0), w.data_ptr(), L, D, float(eps), _p(cos), _p(sin), int(per_head), _stream())
    return x


def qk_rmsnorm_rope_(q, k, wq, wk, eps=1e-6, cos=None, sin=None, per_head=False):
    """rmsnorm_rope_ of the q and the k column block of one fused buffer in a single launch (same row stride)."""
    _chk(q, bf16, "q"), _chk(k, bf16, "k"), _chk(wq, f32, "wq"), _chk(wk, f32, "wk")
    L, D = q.shape
    assert k.shape == q.shape and q.stride(1) == 1 and k.stride(1) == 1 and q.stride(0) == k.stride(0)
    assert wq.numel() == (128 if per_head else D) and wk.numel() == wq.numel()
    if cos is not None:
        _chk(cos, f32, "cos"), _chk(sin, f32, "sin")
        assert cos.shape == (L, 128) and sin.shape == (L, 128) and cos.is_contiguous() and sin.is_contiguous()
    _lib.call("b200_qk_rmsnorm_rope", q.data_ptr(), k.data_ptr(), q.stride(0), wq.data_ptr(), wk.data_ptr(), L, D, float(eps), _p(cos), _p(sin),
              int(per_head), _stream())
    return q, k


def attention(q, k, v, num_heads, out=None, scale=None, nseq=1):
    """q [Lq, H*128], k/v [Lk, H*128] bf16 (row-strided views allowed) -> [Lq, H*128] bf16.
    nseq > 1: nseq equally long sequences stacked along the rows (q [nseq*Lq, .], k/v [nseq*Lk, .]), each attending to its own keys."""
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        _chk(t, bf16, n)
        assert t.dim() == 2 and t.stride(1) == 1 and t.shape[1] == num_heads * 128
    assert q.shape[0] % nseq == 0 and k.shape[0] % nseq == 0 and v.shape[0] == k.shape[0]
    Lq, Lk = q.shape[0] // nseq, k.shape[0] // nseq
    if out is None:
        out = torch.empty(q.shape[0], num_heads * 128, device=q.device, dtype=bf16)
    scale = 1.0 / math.sqrt(128) if scale is None else scale
    if nseq == 1:
        fn = lambda: _lib.call("b200_attention_d128", q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), Lq, Lk, num_heads,   # noqa: E731
                               q.stride(0), k.stride(0), v.stride(0), out.stride(0), float(scale), _stream())
    else:
        fn = lambda: _lib.call("b200_attention_d128_batched", q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), nseq, Lq, Lk,   # noqa: E731
                               num_heads, q.stride(0), k.stride(0), v.stride(0), out.stride(0), float(scale), _stream())
    _timed("attention" if Lq == Lk else "cross_attention", 4.0 * nseq * Lq * Lk * num_heads * 128, fn)
    return out


def cast_bf16(x):
    _chk(x, f32, "x")
    x = x.contiguous()
    y = torch.empty(x.shape, device=x.device, dtype=bf16)
    _lib.call("b200_cast_f32_bf16", x.data_ptr(), y.data_ptr(), x.numel(), _stream())
    return y


def patch_embed(x, y, w, bias, D, patch=2, out=None):
    """x [C0,T,H,W] fp32, y [C1,T,H,W] fp32 or None, w [D, (C0+C1)*patch^2] fp32 -> [L, D] fp32."""
    _chk(x, f32, "x"), _chk(w, f32, "w"), _chk(bias, f32, "bias")
    C0, T, H, W = x.shape
    C1 = 0 if y is None else y.shape[0]
    assert x.is_contiguous() and (y is None or y.is_contiguous()) and w.is_contiguous()
    L = T * (H // patch) * (W // patch)
    if out is None:
        out = torch.empty(L, D, device=x.device, dtype=f32)
    assert out.shape == (L, D) and out.is_contiguous() and out.dtype == f32
    _lib.call("b200_patch_embed", x.data_ptr(), C0, _p(y), C1, w.data_ptr(), bias.data_ptr(), out.data_ptr(), T, H, W, D,
              int(patch), _stream())
    return out


def unpatchify(y, C, T, H, W, patch=2, c_major=False):
    _chk(y, f32, "y")
    assert y.is_contiguous() and y.shape == (T * (H // patch) * (W // patch), patch * patch * C)
    out = torch.empty(C, T, H, W, device=y.device, dtype=f32)
    _lib.call("b200_unpatchify", y.data_ptr(), out.data_ptr(), C, T, H, W, int(patch), int(c_major), _stream())
    return out


def gemv(x, w, b, silu_in=False, silu_out=False):
    _chk(x, f32, "x"), _chk(w, f32, "w"), _chk(b, f32, "b")
    N, K = w.shape
    assert x.numel() == K and w.is_contiguous()
    out = torch.empty(N, device=x.device, dtype=f32)
    _lib.call("b200_gemv_f32", x.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), N, K, int(silu_in), int(silu_out),
              _stream())
    return out


def sinusoid(t, dim, device):
    """t: a number, or a 1-element fp32 CUDA tensor (read on the device: nothing is baked into a captured graph)."""
    out = torch.empty(dim, device=device, dtype=f32)
    if torch.is_tensor(t) and t.is_cuda:
        _chk(t, f32, "t")
        _lib.call("b200_sinusoid_dev", t.data_ptr(), out.data_ptr(), dim, _stream())
    else:
        _lib.call("b200_sinusoid", float(t), out.data_ptr(), dim, _stream())
    return out


def col_mean(x):
    _chk(x, f32, "x")
    assert x.dim() == 2 and x.is_contiguous()
    out = torch.empty(x.shape[1], device=x.device, dtype=f32)
    _lib.call("b200_col_mean_f32", x.data_ptr(), out.data_ptr(), x.shape[0], x.shape[1], _stream())
    return out


def add_vec(a, b):
    _chk(a, f32, "a"), _chk(b, f32, "b")
    out = torch.empty_like(a)
    _lib.call("b200_add_vec", a.data_ptr(), b.data_ptr(), out.data_ptr(), a.numel(), b.numel(), _stream())
    return out


def _chk_vec4(*tensors):
    """The step kernels use 128-bit loads: every operand must start on a 16-byte boundary (a view at an odd float offset does not)."""
    for t in tensors:
        if t is not None and t.data_ptr() % 16:
            raise _lib.B200Error("cfg step: operand not 16-byte aligned (a sliced view?); pass an aligned / contiguous tensor")


def cfg_euler_step_(lat, cond, uncond, guide, dt, pred_out=None, cfg_star=False):
    """lat -= dt * (u + g (c - u)); cfg_star=True applies the CFG-Zero* rescale of u (any2video.py:1706-1714)."""
    _chk(lat, f32, "lat"), _chk(cond, f32, "cond")
    _chk_vec4(lat, cond, uncond, pred_out)
    assert lat.is_contiguous() and cond.is_contiguous() and (uncond is None or uncond.is_contiguous())
    dots = torch.empty(CFG_DOTS_FLOATS, device=lat.device, dtype=f32) if cfg_star else None
    if torch.is_tensor(dt):              # fp32 CUDA tensor [guide, dt]: read on the device (whole-step CUDA graph); `guide` is ignored
        _chk(dt, f32, "guide_dt")
        assert dt.numel() == 2 and dt.is_contiguous()
        _lib.call("b200_cfg_euler_step_dev", lat.data_ptr(), cond.data_ptr(), _p(uncond), dt.data_ptr(), _p(pred_out), _p(dots), lat.numel(), _stream())
        return lat
    _lib.call("b200_cfg_euler_step", lat.data_ptr(), cond.data_ptr(), _p(uncond), float(guide), float(dt), _p(pred_out),
              _p(dots), lat.numel(), _stream())
    return lat


def cfg_unipc_step_(lat, cond, uncond, guide, x_last, m0, m1, coef, cfg_star=False):
    """One FlowUniPCMultistepScheduler.step fused with the CFG combine (pipeline.UniPCSchedule.coefficients gives `coef`):
    lat <- x_next, x_last <- corrected sample, m1 <- x0 of this step (the caller swaps m0/m1)."""
    for t_, n_ in ((lat, "lat"), (cond, "cond"), (x_last, "x_last"), (m0, "m0"), (m1, "m1")):
        _chk(t_, f32, n_)
        assert t_.is_contiguous() and t_.numel() == lat.numel()
    assert uncond is None or (uncond.is_contiguous() and uncond.numel() == lat.numel())
    _chk_vec4(lat, cond, uncond, x_last, m0, m1)
    dots = torch.empty(CFG_DOTS_FLOATS, device=lat.device, dtype=f32) if cfg_star else None
    if torch.is_tensor(coef):            # fp32 CUDA tensor [guide, sigma, ca, cb, cc, cd, pp, pq, pr, use_corrector]: read on the device
        _chk(coef, f32, "coef")
        assert coef.numel() == 10 and coef.is_contiguous()
        _lib.call("b200_cfg_unipc_step_dev", lat.data_ptr(), cond.data_ptr(), _p(uncond), x_last.data_ptr(), m0.data_ptr(), m1.data_ptr(),
                  coef.data_ptr(), _p(dots), lat.numel(), _stream())
        return lat
    c = (ctypes.c_float * 8)(coef["sigma"], coef["ca"], coef["cb"], coef["cc"], coef["cd"], coef["pp"], coef["pq"], coef["pr"])
    _lib.call("b200_cfg_unipc_step", lat.data_ptr(), cond.data_ptr(), _p(uncond), float(guide), x_last.data_ptr(), m0.data_ptr(),
              m1.data_ptr(), ctypes.addressof(c), int(coef["use_corrector"]), _p(dots), lat.numel(), _stream())
    return lat
