"""GPU parity tests of the individual kernels, called through the C ABI (wan2gp_b200.ops -> ctypes).
Each kernel is compared with a plain PyTorch fp32/fp64 evaluation of the same op on the same
bf16-rounded inputs.  Tolerances (stated per SURVEY.md section 7/H1):
  * fp32-output kernels: rel-L2 <= 1e-3 (observed ~1e-6 .. 1e-4)
  * bf16-output kernels: rel-L2 <= 4e-3 (one bf16 rounding of the result is 2^-9 relative)
"""
import math

import pytest
import torch

from tests.helpers import rel_l2

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(120)]

bf16, f32 = torch.bfloat16, torch.float32
TOL_F32, TOL_BF16 = 1e-3, 4e-3


@pytest.fixture(scope="module")
def ops():
    from wan2gp_b200 import ops as o
    return o


def _randn(*shape, seed=0, scale=1.0, dtype=f32):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, generator=g, device="cuda", dtype=f32) * scale).to(dtype)


@pytest.mark.parametrize("M,N,K", [(256, 1472, 384), (300, 1472, 3584), (130, 800, 128), (640, 1184, 256)])
def test_gemm_accumulate_with_n_tail(ops, M, N, K):
    """fp32 residual-stream accumulate (`x += gate * (a @ b^T + bias)`) when N is not a multiple of the 256-wide tile (byT5: 1472): the
    TMA reduce-add epilogue must not let the chunks beyond N touch its staging buffers (round 2, call 26: the last 32 valid columns raced with
    the reduce-store still reading them).  Repeated launches, every column checked."""
    a, b = _randn(M, K, seed=1, dtype=bf16), _randn(N, K, seed=2, scale=K ** -0.5, dtype=bf16)
    bias, gate = _randn(N, seed=3), _randn(N, seed=4)
    lin = (a.double() @ b.double().t() + bias.double()) * gate.double()
    x0 = _randn(M, N, seed=5)
    for rep in range(4):
        x = x0.clone()
        ops.gemm(a, b, out=x, bias=bias, gate=gate, accumulate=True)
        err = (x.double() - (x0.double() + lin)).abs()
        assert float(err.max()) < 1e-3 * float(lin.abs().max()), (rep, int(err.argmax()) % N)
    assert rel_l2(x[:, N - 32:], (x0.double() + lin)[:, N - 32:]) < TOL_F32


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (128, 256, 512), (300, 512, 320), (1000, 768, 1536),
                                   (130, 64, 128), (4096, 1536, 1536), (77, 96, 72), (257, 192, 200), (64, 1560, 384)])
def test_gemm_bf16_out(ops, M, N, K):
    a, b = _randn(M, K, seed=1, dtype=bf16), _randn(N, K, seed=2, scale=K ** -0.5, dtype=bf16)
    bias = _randn(N, seed=3)
    out = ops.gemm(a, b, bias=bias)
    ref = a.double() @ b.double().t() + bias.double()
    assert out.dtype == bf16 and rel_l2(out, ref) < TOL_BF16


def test_gemm_epilogues(ops):
    M, N, K = 520, 512, 256
    a, b = _randn(M, K, seed=1, dtype=bf16), _randn(N, K, seed=2, scale=K ** -0.5, dtype=bf16)
    bias, gate = _randn(N, seed=3), _randn(N, seed=4)
    lin = a.double() @ b.double().t() + bias.double()
    # fp32 out
    assert rel_l2(ops.gemm(a, b, bias=bias, out_dtype=f32), lin) < TOL_F32
    # no bias
    assert rel_l2(ops.gemm(a, b, out_dtype=f32), a.double() @ b.double().t()) < TOL_F32
    # GELU(tanh) -> bf16      (model.py:551-553 ffn.1)
    ref = torch.nn.functional.gelu(lin.float(), approximate="tanh")
    assert rel_l2(ops.gemm(a, b, bias=bias, act=1), ref) < TOL_BF16
    # gated residual accumulate in fp32   (model.py:659 x.addcmul_(y, e[2]))
    x0 = _randn(M, N, seed=5)
    x = x0.clone()
    ops.gemm(a, b, out=x, bias=bias, gate=gate, accumulate=True)
    assert rel_l2(x, x0.double() + lin * gate.double()) < TOL_F32
    # bf16 residual add (VAE ResidualBlock skip)
    r = _randn(M, N, seed=6, dtype=bf16)
    assert rel_l2(ops.gemm(a, b, bias=bias, residual=r), lin + r.double()) < TOL_BF16
    # strided A / out views (fused qkv buffers)
    big = _randn(M, 3 * K, seed=7, dtype=bf16)
    outbig = torch.zeros(M, 2 * N, device="cuda", dtype=bf16)
    ops.gemm(big[:, K:2 * K], b, out=outbig[:, N:], bias=bias)
    assert rel_l2(outbig[:, N:], big[:, K:2 * K].double() @ b.double().t() + bias.double()) < TOL_BF16
    assert outbig[:, :N].abs().max() == 0


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (200, 128, 200), (512, 384, 1560)])
def test_gemm_b_mn_major(ops, M, N, K):
    """B given as [K, N] (N contiguous): exercises the MN-major UMMA descriptor used for V in attention."""
    a, b = _randn(M, K, seed=1, dtype=bf16), _randn(K, N, seed=2, scale=K ** -0.5, dtype=bf16)
    out = ops.gemm(a, b, out_dtype=f32, b_mn_major=True)
    assert rel_l2(out, a.double() @ b.double()) < TOL_F32


def test_ln_modulate(ops):
    for L, D in [(37, 256), (300, 1536), (130, 5120)]:
        x = _randn(L, D, seed=1, scale=3.0) + 0.5
        sh, sc = _randn(D, seed=2), _randn(D, seed=3, scale=0.3)
        xn = torch.nn.functional.layer_norm(x.double(), (D,), eps=1e-6)
        assert rel_l2(ops.ln_modulate(x, sh, sc), xn * (1 + sc.double()) + sh.double()) < TOL_BF16
        assert rel_l2(ops.ln_modulate(x, sh, sc, affine=True), xn * sc.double() + sh.double()) < TOL_BF16


def test_rmsnorm_rope(ops):
    from oracle import wan_oracle
    for (T, H, W), heads in [((2, 4, 6), 2), ((3, 8, 12), 12), ((1, 4, 4), 40)]:
        cos, sin = wan_oracle.rope_tables((T, H, W))
        L, D = cos.shape[0], heads * 128
        buf = _randn(L, 3 * D, seed=1, dtype=bf16)          # fused qkv layout, normalise the k slice in place
        w = 0.3 * (1 + _randn(D, seed=2, scale=0.1))
        x = buf[:, D:2 * D]
        ref = wan_oracle.rms_norm_full(x.float().cpu(), w.cpu(), 1e-6)
        ref_rope = wan_oracle.apply_rope(ref.reshape(L, heads, 128), cos, sin).reshape(L, D)
        keep = buf.clone()
        ops.rmsnorm_rope_(x, w, 1e-6, cos.cuda(), sin.cuda())
        assert rel_l2(buf[:, D:2 * D].cpu(), ref_rope) < TOL_BF16
        assert torch.equal(buf[:, :D], keep[:, :D]) and torch.equal(buf[:, 2 * D:], keep[:, 2 * D:])
        # q and k of the fused buffer in ONE launch (two weights): bit-identical to the two single-segment launches
        w2 = 0.3 * (1 + _randn(D, seed=9, scale=0.1))
        one, two = keep.clone(), keep.clone()
        ops.rmsnorm_rope_(one[:, :D], w, 1e-6, cos.cuda(), sin.cuda())
        ops.rmsnorm_rope_(one[:, D:2 * D], w2, 1e-6, cos.cuda(), sin.cuda())
        ops.qk_rmsnorm_rope_(two[:, :D], two[:, D:2 * D], w, w2, 1e-6, cos.cuda(), sin.cuda())
        assert torch.equal(one, two)
        y = keep[:, :D].contiguous()
        ops.rmsnorm_rope_(y, w, 1e-6)                        # no RoPE (cross-attention)
        assert rel_l2(y.cpu(), wan_oracle.rms_norm_full(keep[:, :D].float().cpu(), w.cpu(), 1e-6)) < TOL_BF16


@pytest.mark.parametrize("L,heads", [(4099, 40), (2048, 12), (2051, 1)])
def test_rmsnorm_rope_pipelined_kernel_equals_row_kernel(ops, L, heads):
    """L >= 2048 full-width rows take the cp.async-pipelined kernel (8 (row, segment) items per CTA through a 3-stage shared-memory ring;
    4099 rows x 2 segments leaves a last CTA with 6 items): bit-identical to the one-CTA-per-row kernel (which chunks below 2048 rows still
    take) -- same fp32 operations in the same order -- for the single-segment call, the fused q|k call and the call without RoPE; the v
    columns of the fused buffer stay untouched."""
    D = heads * 128
    g = torch.Generator(device="cuda").manual_seed(L)
    cos, sin = torch.randn(L, 128, device="cuda", generator=g), torch.randn(L, 128, device="cuda", generator=g)
    buf = _randn(L, 3 * D, seed=1, dtype=bf16)
    wq, wk = 0.3 * (1 + _randn(D, seed=2, scale=0.1)), 0.3 * (1 + _randn(D, seed=9, scale=0.1))
    want = buf.clone()
    for r0 in range(0, L, 1500):                             # chunks below the threshold: the row kernel
        r1 = min(L, r0 + 1500)
        ops.qk_rmsnorm_rope_(want[r0:r1, :D], want[r0:r1, D:2 * D], wq, wk, 1e-6, cos[r0:r1].contiguous(), sin[r0:r1].contiguous())
    got = buf.clone()
    ops.qk_rmsnorm_rope_(got[:, :D], got[:, D:2 * D], wq, wk, 1e-6, cos, sin)
    assert torch.equal(got, want) and torch.equal(got[:, 2 * D:], buf[:, 2 * D:]) and not torch.equal(got[:, :D], buf[:, :D])
    one = buf.clone()
    ops.rmsnorm_rope_(one[:, D:2 * D], wk, 1e-6, cos, sin)
    assert torch.equal(one[:, D:2 * D], want[:, D:2 * D]) and torch.equal(one[:, :D], buf[:, :D])
    nr, nr_want = buf[:, :D].contiguous(), buf[:, :D].contiguous()
    ops.rmsnorm_rope_(nr, wq, 1e-6)
    for r0 in range(0, L, 1500):
        ops.rmsnorm_rope_(nr_want[r0:min(L, r0 + 1500)], wq, 1e-6)
    assert torch.equal(nr, nr_want)
    # and against fp64 on a sample of rows
    rows = torch.tensor([0, 1, 7, 8, L // 2, L - 7, L - 1], device="cuda")
    x = buf[rows, :D].double()
    y = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-6) * wq.double()
    y = y.reshape(len(rows), heads, 64, 2)
    c, s_ = cos[rows].double().reshape(len(rows), 1, 64, 2), sin[rows].double().reshape(len(rows), 1, 64, 2)
    ref = torch.stack([y[..., 0] * c[..., 0] - y[..., 1] * s_[..., 0], y[..., 1] * c[..., 1] + y[..., 0] * s_[..., 1]], -1).reshape(len(rows), D)
    assert rel_l2(got[rows, :D], ref) < TOL_BF16


# Lq >= 1024 takes the CTA-pair kernel (attn2_sm100.cuh: 512 query rows per cluster): full / ragged pair blocks, a second CTA that is
# entirely out of range (1100 = 2 x 512 + 76), partial last K/V tile, cross-attention length 512
@pytest.mark.parametrize("Lq,Lk,H", [(128, 128, 1), (128, 256, 2), (200, 333, 3), (1000, 1000, 2), (300, 512, 4), (64, 80, 1),
                                     (1024, 1024, 2), (2000, 1333, 3), (1100, 512, 2), (1536, 200, 1), (3510, 3510, 2)])
def test_attention(ops, Lq, Lk, H):
    D = H * 128
    q, k, v = (_randn(L, D, seed=s, dtype=bf16) for L, s in ((Lq, 1), (Lk, 2), (Lk, 3)))
    out = ops.attention(q, k, v, H)
    qh, kh, vh = (t.double().reshape(-1, H, 128).permute(1, 0, 2) for t in (q, k, v))
    ref = torch.softmax(qh @ kh.transpose(1, 2) / math.sqrt(128), -1) @ vh
    assert rel_l2(out, ref.permute(1, 0, 2).reshape(Lq, D)) < TOL_BF16


def test_attention_strided_and_peaky(ops):
    """q/k/v as column slices of one fused [L, 3D] buffer; large logits exercise the lazy-rescale path."""
    L, H = 700, 2
    D = H * 128
    buf = _randn(L, 3 * D, seed=4, dtype=bf16)
    buf[:, :2 * D] *= 3.0
    out = ops.attention(buf[:, :D], buf[:, D:2 * D], buf[:, 2 * D:], H)
    qh, kh, vh = (buf[:, i * D:(i + 1) * D].double().reshape(L, H, 128).permute(1, 0, 2) for i in range(3))
    ref = torch.softmax(qh @ kh.transpose(1, 2) / math.sqrt(128), -1) @ vh
    assert rel_l2(out, ref.permute(1, 0, 2).reshape(L, D)) < TOL_BF16


@pytest.mark.parametrize("Lq,Lk,H,boost", [(700, 1500, 2, 60), (700, 1500, 2, -132), (640, 900, 1, -133), (130, 257, 1, 60)])
def test_attention_scores_outrun_first_tile(ops, Lq, Lk, H, boost):
    """The default kernel (csrc/attn6_sm100.cuh) fixes each row's reference maximum on K/V tile 0 and never rescales; rows whose later
    scores exceed it by far more than fp32 can hold must take the exact three-pass path.  boost > 1: every key beyond the first 200 is
    scaled up (row sums overflow); boost < 0: ONE key, at a position whose exponential runs on the FMA pipe (132) or on MUFU (133), is 80x
    larger than the rest (a polynomial exp2 argument beyond 127 is garbage, not +inf -- the guard must catch it)."""
    D = H * 128
    q, k, v = (_randn(L, D, seed=s, dtype=bf16) for L, s in ((Lq, 1), (Lk, 2), (Lk, 3)))
    if boost > 1:
        k[200:] *= boost
    else:
        k[-boost] *= 80
    out = ops.attention(q, k, v, H)
    qh, kh, vh = (t.double().reshape(-1, H, 128).permute(1, 0, 2) for t in (q, k, v))
    ref = torch.softmax(qh @ kh.transpose(1, 2) / math.sqrt(128), -1) @ vh
    assert torch.isfinite(out).all()
    assert rel_l2(out, ref.permute(1, 0, 2).reshape(Lq, D)) < TOL_BF16


def test_patch_embed_unpatchify_gemv(ops):
    from oracle import wan_oracle
    C0, C1, T, H, W, D = 16, 20, 3, 8, 12, 256
    x, y = _randn(C0, T, H, W, seed=1), _randn(C1, T, H, W, seed=2)
    w, b = _randn(D, (C0 + C1) * 4, seed=3, scale=0.1), _randn(D, seed=4)
    sd = {"patch_embedding.weight": w.cpu().reshape(D, C0 + C1, 1, 2, 2), "patch_embedding.bias": b.cpu()}
    assert rel_l2(ops.patch_embed(x, y, w, b, D).cpu(), wan_oracle.patch_embed(sd, torch.cat([x, y]).cpu())) < 1e-5
    sd = {"patch_embedding.weight": w[:, :64].cpu().reshape(D, C0, 1, 2, 2), "patch_embedding.bias": b.cpu()}
    assert rel_l2(ops.patch_embed(x, None, w[:, :64].contiguous(), b, D).cpu(), wan_oracle.patch_embed(sd, x.cpu())) < 1e-5
    L = T * (H // 2) * (W // 2)
    yv = _randn(L, 64, seed=5)
    ref = torch.einsum("fhwpqrc->cfphqwr", yv.cpu().reshape(T, H // 2, W // 2, 1, 2, 2, 16)).reshape(16, T, H, W)
    assert torch.equal(ops.unpatchify(yv, 16, T, H, W).cpu(), ref)
    xv, wv, bv = _randn(256, seed=6), _randn(512, 256, seed=7, scale=0.1), _randn(512, seed=8)
    ref = torch.nn.functional.silu(torch.nn.functional.silu(xv) @ wv.t() + bv)
    assert rel_l2(ops.gemv(xv, wv, bv, silu_in=True, silu_out=True), ref) < 1e-5
    assert rel_l2(ops.sinusoid(500.0, 256, "cuda").cpu(), wan_oracle.sinusoidal_embedding(256, torch.tensor([500.0]))[0]) < 1e-5


def test_cfg_euler(ops):
    from oracle import wan_oracle
    lat, c, u = (_randn(1, 16, 3, 8, 12, seed=s) for s in (1, 2, 3))
    ref = wan_oracle.euler_step(lat.cpu(), wan_oracle.cfg_combine(c.cpu(), u.cpu(), 4.0), 0.9, 0.85)
    lat0 = lat.clone()
    ops.cfg_euler_step_(lat, c, u, 4.0, 0.05)
    assert rel_l2(lat.cpu(), ref) < 1e-6
    # CFG-Zero* (any2video.py:1706-1714)
    ref = wan_oracle.euler_step(lat0.cpu(), wan_oracle.cfg_combine(c.cpu(), u.cpu(), 4.0, cfg_star=True, step_no=3), 0.9, 0.85)
    ops.cfg_euler_step_(lat0, c, u, 4.0, 0.05, cfg_star=True)
    assert rel_l2(lat0.cpu(), ref) < 1e-5


@pytest.mark.parametrize("nseq,Lq,Lk,H", [(2, 300, 300, 2), (2, 1000, 512, 3), (3, 130, 77, 1), (2, 3510, 3510, 2)])
def test_attention_batched_equals_separate(ops, nseq, Lq, Lk, H):
    """nseq sequences stacked along the rows in ONE launch (CFG cond / uncond) == one launch per sequence, bit for bit: ragged last Q
    tiles (rows of the next sequence are loaded but never stored) and ragged last K/V tiles (rows of the next sequence are masked)."""
    D = H * 128
    q, k, v = (_randn(nseq * n, D, seed=s, dtype=bf16) for n, s in ((Lq, 1), (Lk, 2), (Lk, 3)))
    both = ops.attention(q, k, v, H, nseq=nseq)
    for z in range(nseq):
        one = ops.attention(q[z * Lq:(z + 1) * Lq], k[z * Lk:(z + 1) * Lk], v[z * Lk:(z + 1) * Lk], H)
        assert torch.equal(both[z * Lq:(z + 1) * Lq], one), z


def test_cfg_step_rejects_misaligned_views(ops):
    """The fused CFG + scheduler kernels use 128-bit loads: an operand that is a view at an odd float offset must be refused up front
    (round 2: a payload[1:] view of the CFG-pair exchange buffer reached the kernel and raised 'misaligned address' on the GPU)."""
    from wan2gp_b200 import _lib
    buf = _randn(1 + 4 * 96, seed=1)
    lat, c = _randn(4 * 96, seed=2), _randn(4 * 96, seed=3)
    with pytest.raises(_lib.B200Error):
        ops.cfg_euler_step_(lat, buf[1:], c, 4.0, 0.05)
    ops.cfg_euler_step_(lat, buf[4:4 + 384], c, 4.0, 0.05)          # a 16-byte aligned view is fine
