"""CPU oracle for the Wan DiT denoise forward (TEST INFRASTRUCTURE ONLY).

A plain-PyTorch fp32 restatement of the reference algorithm for hot-path rows
W0-W12 of SURVEY.md section 8a, written against the reference state-dict names.  Only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this; the product path (wan2gp_b200/) never does.

Pinned: the reference has no golden vectors for this path, so the
restatement is pinned against the reference's own modules executed in this
container (oracle/gen_golden.py -> tests/golden/*.npz, and
tests/test_oracle_vs_reference.py which runs when /root/reference is present).

`emulate_bf16=True` inserts bf16 roundings at exactly the points where the CUDA
path stores bf16 (GEMM operands / outputs, attention probabilities), so kernel
parity can be checked tightly; `False` is the reference's fp32 CPU arithmetic.
"""
import math

import torch
import torch.nn.functional as F


def _q(t, emulate):
    return t.to(torch.bfloat16).to(torch.float32) if emulate else t


# ---- W0: RoPE tables -- posemb_layers.py:492-525 (get_rotary_pos_embed), 346-431, 434-482
def rope_tables(latent_thw, rope_dims=(44, 42, 42), theta=10000.0):
    """cos, sin: fp32 [L, 128]; token order (t, h', w') row-major, h'=H/2, w'=W/2 (patch 1,2,2)."""
    T, H, W = latent_thw
    sizes = (T, H // 2, W // 2)
    grids = torch.meshgrid(*[torch.arange(n, dtype=torch.float32) for n in sizes], indexing="ij")
    cos, sin = [], []
    for d, g in zip(rope_dims, grids):
        # posemb_layers.py:470-477: freqs = 1/theta^(arange(0,d,2)/d); outer(pos, freqs); repeat_interleave(2)
        inv = 1.0 / (theta ** (torch.arange(0, d, 2, dtype=torch.float32)[: d // 2] / d))
        ang = torch.outer(g.reshape(-1), inv)
        cos.append(ang.cos().repeat_interleave(2, dim=1))
        sin.append(ang.sin().repeat_interleave(2, dim=1))
    return torch.cat(cos, 1), torch.cat(sin, 1)


# ---- W6: interleaved-pair rotation -- posemb_layers.py:251-259
def apply_rope(x, cos, sin):
    """x [L, H, 128] fp32; o[2j] = x[2j]c - x[2j+1]s ; o[2j+1] = x[2j+1]c + x[2j]s."""
    xv = x.reshape(*x.shape[:-1], -1, 2)
    c = cos.reshape(cos.shape[0], 1, -1, 2)
    s = sin.reshape(sin.shape[0], 1, -1, 2)
    o0 = xv[..., 0] * c[..., 0] - xv[..., 1] * s[..., 0]
    o1 = xv[..., 1] * c[..., 1] + xv[..., 0] * s[..., 1]
    return torch.stack([o0, o1], -1).reshape(x.shape)


# ---- W1: model.py:32-42 sinusoidal_embedding_1d
def sinusoidal_embedding(dim, t):
    half = dim // 2
    s = torch.outer(t.float(), torch.pow(10000.0, -torch.arange(half, dtype=torch.float32) / half))
    return torch.cat([s.cos(), s.sin()], 1)


def layer_norm(x, eps):            # WanLayerNorm without affine, model.py:194-212
    mu = x.mean(-1, keepdim=True)
    var = (x - mu).pow(2).mean(-1, keepdim=True)
    return (x - mu) * torch.rsqrt(var + eps)


# Set by tests that pin the restatement against the reference's *fp32* run.  In WanRMSNorm.forward
# (model.py:165-172) `y = x.float()` ALIASES x when x is already fp32, so `y.pow_(2)` squares x in
# place before `x *= rsqrt(mean)`: the fp32 CPU path computes x^2 * rsqrt(mean(x^2)+eps) * w.  The
# production path (bf16 activations, or any dtype != fp32) gets a copy and computes the true RMSNorm.
# We implement the production semantics; the quirk is reproduced here only to prove the restatement
# matches the reference bit-for-bit-ish in fp32 as well (see DESIGN.md "oracle pinning").
FP32_ALIAS_QUIRK = False


def rms_norm_full(x, w, eps):      # WanRMSNorm over the FULL dim, model.py:152-175
    r = torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps)
    if FP32_ALIAS_QUIRK:
        return x * x * r * w
    return x * r * w


def gelu_tanh(x):
    return F.gelu(x, approximate="tanh")


def attention(q, k, v, emulate):
    """q [Lq,H,128], k/v [Lk,H,128] -> [Lq,H,128]; softmax(q k^T / sqrt(128)) v, no mask
    (shared/attention.py:208-225 sdpa_wrapper)."""
    q, k, v = (t.permute(1, 0, 2) for t in (q, k, v))
    s = torch.matmul(q, k.transpose(1, 2)) / math.sqrt(q.shape[-1])
    m = s.max(-1, keepdim=True).values
    p = torch.exp(s - m)
    l = p.sum(-1, keepdim=True)
    o = torch.matmul(_q(p, emulate), v) / l   # flash kernels feed bf16 P to the PV MMA, fp32 row sum
    return o.permute(1, 0, 2)


def linear(x, sd, name, emulate):
    w = _q(sd[name + ".weight"].float(), emulate)
    return x @ w.t() + sd[name + ".bias"].float()


# ---- W1/W2 conditioning -- model.py:1815-1818, 1856
def time_conditioning(sd, cfg, t):
    e = sinusoidal_embedding(cfg["freq_dim"], t.flatten())
    e = F.silu(e @ sd["time_embedding.0.weight"].float().t() + sd["time_embedding.0.bias"].float())
    e = e @ sd["time_embedding.2.weight"].float().t() + sd["time_embedding.2.bias"].float()   # [1, D]
    e0 = F.silu(e) @ sd["time_projection.1.weight"].float().t() + sd["time_projection.1.bias"].float()
    return e, e0.reshape(-1, 6, cfg["dim"])


def text_embedding(sd, ctx, emulate):
    h = _q(gelu_tanh(linear(_q(ctx, emulate), sd, "text_embedding.0", emulate)), emulate)
    return _q(linear(h, sd, "text_embedding.2", emulate), emulate)


# ---- W2 patch embed -- model.py:1131-1132, 1631, 1731 (Conv3d k=s=(1,2,2) == per-token GEMM)
def patch_embed(sd, x):
    """x [Cin,T,H,W] fp32 -> [L, D]; K ordered (c, ph, pw); tokens (t, h', w')."""
    Cin, T, H, W = x.shape
    p = x.reshape(Cin, T, H // 2, 2, W // 2, 2).permute(1, 2, 4, 0, 3, 5).reshape(T * (H // 2) * (W // 2), Cin * 4)
    w = sd["patch_embedding.weight"].float().reshape(-1, Cin * 4)
    return p @ w.t() + sd["patch_embedding.bias"].float()


# ---- W3-W10: one WanAttentionBlock -- model.py:631-711
def block_forward(sd, cfg, i, x, e0, ctx, cos, sin, emulate, taps=None, rows=None):
    """rows (LongTensor or None): evaluate the block only for these token rows -- keys / values still come from ALL rows of x, every
    other operation of the block is row-wise -- and return [len(rows), D].  Used by the production-shape parity tests (L = 75 600),
    where the full L x L attention of the restatement would not fit."""
    D, H, eps = cfg["dim"], cfg["num_heads"], cfg["eps"]
    p = f"blocks.{i}."
    m = (sd[p + "modulation"].float() + e0).reshape(6, D)               # model.py:632
    # self-attention (model.py:634-660)
    a = _q(layer_norm(x, eps) * (1 + m[1]) + m[0], emulate)
    q = _q(linear(a if rows is None else a[rows], sd, p + "self_attn.q", emulate), emulate)
    k = _q(linear(a, sd, p + "self_attn.k", emulate), emulate)
    v = _q(linear(a, sd, p + "self_attn.v", emulate), emulate)
    q = rms_norm_full(q, sd[p + "self_attn.norm_q.weight"].float(), eps).reshape(-1, H, D // H)
    k = rms_norm_full(k, sd[p + "self_attn.norm_k.weight"].float(), eps).reshape(-1, H, D // H)
    q = _q(apply_rope(q, cos if rows is None else cos[rows], sin if rows is None else sin[rows]), emulate)
    k = _q(apply_rope(k, cos, sin), emulate)
    o = _q(attention(q, k, v.reshape(-1, H, D // H), emulate).reshape(-1, D), emulate)
    if taps is not None:
        taps[f"b{i}.a"], taps[f"b{i}.q"], taps[f"b{i}.k"], taps[f"b{i}.attn"] = a, q.reshape(-1, D), k.reshape(-1, D), o
    del a, q, k, v
    if rows is not None:
        x = x[rows]
    x = x + linear(o, sd, p + "self_attn.o", emulate) * m[2]
    # text cross-attention (model.py:663-668, 245-265, 444)
    c = F.layer_norm(x, (D,), sd[p + "norm3.weight"].float(), sd[p + "norm3.bias"].float(), eps)
    c = _q(c, emulate)
    q = _q(linear(c, sd, p + "cross_attn.q", emulate), emulate)
    q = _q(rms_norm_full(q, sd[p + "cross_attn.norm_q.weight"].float(), eps), emulate).reshape(-1, H, D // H)
    k = _q(linear(ctx, sd, p + "cross_attn.k", emulate), emulate)
    k = _q(rms_norm_full(k, sd[p + "cross_attn.norm_k.weight"].float(), eps), emulate).reshape(-1, H, D // H)
    v = _q(linear(ctx, sd, p + "cross_attn.v", emulate), emulate).reshape(-1, H, D // H)
    o = _q(attention(q, k, v, emulate).reshape(-1, D), emulate)
    x = x + linear(o, sd, p + "cross_attn.o", emulate)
    if taps is not None:
        taps[f"b{i}.x_cross"] = x
    # FFN (model.py:686-711)
    f = _q(layer_norm(x, eps) * (1 + m[4]) + m[3], emulate)
    h = _q(gelu_tanh(linear(f, sd, p + "ffn.0", emulate)), emulate)
    x = x + linear(h, sd, p + "ffn.2", emulate) * m[5]
    return x


# ---- W11 head + unpatchify -- model.py:847-865, 2100-2126
def head_unpatchify(sd, cfg, x, e, thw):
    D, eps, C = cfg["dim"], cfg["eps"], cfg["out_dim"]
    T, H, W = thw
    h = (sd["head.modulation"].float() + e.reshape(1, 1, D)).reshape(2, D)
    y = layer_norm(x, eps) * (1 + h[1]) + h[0]
    y = y @ sd["head.head.weight"].float().t() + sd["head.head.bias"].float()     # [L, 4*C] order (ph, pw, c)
    y = y.reshape(T, H // 2, W // 2, 1, 2, 2, C)
    y = torch.einsum("fhwpqrc->cfphqwr", y)
    return y.reshape(C, T, H, W)


def wan_forward(sd, cfg, x, t, context, y=None, freqs=None, emulate_bf16=False, taps=None, num_layers=None):
    """Restates WanModel.forward (model.py:1485-2098) for the plain t2v / i2v2_2 path.
    x [B,Cin,T,H,W] fp32, t [1], context [1,text_len,text_dim], y [Cy,T,H,W] or None
    -> [B,16,T,H,W] fp32."""
    B, _, T, H, W = x.shape
    cos, sin = freqs if freqs is not None else rope_tables((T, H, W))
    e, e0 = time_conditioning(sd, cfg, t)
    ctx = text_embedding(sd, context[0].float(), emulate_bf16)
    outs = []
    nl = cfg["num_layers"] if num_layers is None else num_layers
    for b in range(B):
        xb = x[b].float()
        if y is not None:
            xb = torch.cat([xb, y.float()], 0)                              # model.py:1597-1600
        h = patch_embed(sd, xb)
        if taps is not None:
            taps["patch"] = h
        for i in range(nl):
            h = block_forward(sd, cfg, i, h, e0[0], ctx, cos, sin, emulate_bf16, taps)
            if taps is not None:
                taps[f"b{i}.x"] = h
        outs.append(head_unpatchify(sd, cfg, h, e[0], (T, H, W)))
    return torch.stack(outs, 0)


# ---- W12 CFG combine -- any2video.py:1701-1722 (plain CFG and CFG-Zero*), flow-match Euler step
def cfg_combine(cond, uncond, guide_scale, cfg_star=False, step_no=0, cfg_zero_step=-1):
    if cfg_star:
        b = cond.shape[0]
        dot = (cond.reshape(b, -1) * uncond.reshape(b, -1)).sum(1, keepdim=True)
        sq = uncond.reshape(b, -1).pow(2).sum(1, keepdim=True) + 1e-8
        alpha = (dot / sq).reshape(b, 1, 1, 1, 1)
        if step_no > cfg_zero_step:
            uncond = uncond * alpha
    return uncond + guide_scale * (cond - uncond)


def euler_step(latents, noise_pred, sigma, sigma_next):
    """shared/utils/euler_scheduler.py:67-86 -- x <- x + (sigma_next - sigma) * v."""
    return latents + (sigma_next - sigma) * noise_pred


def unipc_step(x, v, x_last, m0, m1, coef):
    """CPU restatement of the fused UniPC update (csrc/elementwise.cuh::cfg_unipc_kernel) for given scalar coefficients
    (wan2gp_b200/pipeline.py::UniPCSchedule.coefficients): returns (x_next, x_corrected, x0).  Pinned end to end against the reference
    FlowUniPCMultistepScheduler.step (shared/utils/fm_solvers_unipc.py:655-740) in tests/test_unipc_cpu.py."""
    x0 = x - coef["sigma"] * v
    xc = coef["ca"] * x_last + coef["cb"] * m0 + coef["cc"] * m1 + coef["cd"] * x0 if coef["use_corrector"] else x
    xn = (coef["pp"] * xc if coef["pp"] != 0 else 0) + coef["pq"] * x0 + (coef["pr"] * m0 if coef["pr"] != 0 else 0)
    return xn, xc, x0
