"""CPU oracle for the Hunyuan Video 1.5 DiT forward (TEST INFRASTRUCTURE ONLY) -- hot-path rows H1-H4 of SURVEY.md
section 8a: models/hyvideo/modules/models.py::HYVideoDiffusionTransformer.forward (:946-1233) for the 'HYVideo-1_5' family
(54 double-stream blocks, pre-split qkv, token refiner, byT5 mapper, cond-type embedding, patch (1,1,1)).

Pinning: the reference blocks hard-cast activations to bf16 (models.py:211, 290, 302, 313), so it cannot run with plain
fp32 weights.  oracle/gen_golden.py runs it with fp32 weights and a forward-pre-hook that casts each Linear's INPUT to the
weight dtype (oracle/refshim.py::hook_linear_input_cast) => fp32 arithmetic with bf16 roundings exactly at the reference's
hard-cast points.  `ref_casts=True` reproduces those roundings; `emulate_bf16=True` additionally rounds where the CUDA path
stores bf16 (GEMM operands / outputs, attention probabilities).
"""
import math

import torch
import torch.nn.functional as F

from oracle.wan_oracle import apply_rope


def _q(t, on):
    return t.to(torch.bfloat16).to(torch.float32) if on else t


def rope_tables_hy(thw, rope_dims=(16, 56, 56), theta=256.0):
    """hunyuan.py:677-725 -> posemb_layers.get_nd_rotary_pos_embed(theta=256): cos/sin fp32 [L, 128]."""
    grids = torch.meshgrid(*[torch.arange(n, dtype=torch.float32) for n in thw], indexing="ij")
    cos, sin = [], []
    for d, g in zip(rope_dims, grids):
        inv = 1.0 / (theta ** (torch.arange(0, d, 2, dtype=torch.float32)[: d // 2] / d))
        ang = torch.outer(g.reshape(-1), inv)
        cos.append(ang.cos().repeat_interleave(2, dim=1))
        sin.append(ang.sin().repeat_interleave(2, dim=1))
    return torch.cat(cos, 1), torch.cat(sin, 1)


def timestep_embedding(t, dim=256, max_period=10000):        # embed_layers.py:110-134
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], -1)


def lin(sd, name, x, em=False):
    return _q(x, em) @ _q(sd[name + ".weight"].float(), em).t() + sd[name + ".bias"].float()


def timestep_embedder(sd, p, t):                              # embed_layers.py:137-174 (fp32, GEMV-sized)
    h = F.silu(timestep_embedding(t) @ sd[p + "mlp.0.weight"].float().t() + sd[p + "mlp.0.bias"].float())
    return h @ sd[p + "mlp.2.weight"].float().t() + sd[p + "mlp.2.bias"].float()


def attention(q, k, v, em):
    """q [Lq,H,128], k/v [Lk,H,128]"""
    q, k, v = (u.permute(1, 0, 2) for u in (q, k, v))
    s = torch.matmul(q, k.transpose(1, 2)) / math.sqrt(q.shape[-1])
    p = torch.exp(s - s.max(-1, keepdim=True).values)
    return (torch.matmul(_q(p, em), v) / p.sum(-1, keepdim=True)).permute(1, 0, 2)


def token_refiner(sd, cfg, txt, t, n_valid, em):
    """SingleTokenRefiner (token_refiner.py:165-237) restricted to the VALID tokens txt [n_valid, text_dim]: valid rows only
    attend valid keys (mask :139-147) and padded rows are zeroed afterwards by reorder_txt_token(zero_feat=True)."""
    D, H = cfg["hidden_size"], cfg["heads_num"]
    p = "txt_in."
    ta = timestep_embedder(sd, p + "t_embedder.", t)[0]
    ctx = txt.mean(0)                                            # masked mean over valid tokens (:221-226)
    ctx = F.silu(ctx @ sd[p + "c_embedder.linear_1.weight"].float().t() + sd[p + "c_embedder.linear_1.bias"].float())
    ctx = ctx @ sd[p + "c_embedder.linear_2.weight"].float().t() + sd[p + "c_embedder.linear_2.bias"].float()
    c = ta + ctx
    x = lin(sd, p + "input_embedder", txt, em)
    for j in range(2):
        b = p + f"individual_token_refiner.blocks.{j}."
        mod = F.silu(c) @ sd[b + "adaLN_modulation.1.weight"].float().t() + sd[b + "adaLN_modulation.1.bias"].float()
        g_msa, g_mlp = mod.chunk(2)
        nx = _q(F.layer_norm(x, (D,), sd[b + "norm1.weight"].float(), sd[b + "norm1.bias"].float(), 1e-6), em)
        qkv = _q(lin(sd, b + "self_attn_qkv", nx, em), em).reshape(-1, 3, H, D // H)
        a = _q(attention(qkv[:, 0], qkv[:, 1], qkv[:, 2], em).reshape(-1, D), em)
        x = x + lin(sd, b + "self_attn_proj", a, em) * g_msa
        nx = _q(F.layer_norm(x, (D,), sd[b + "norm2.weight"].float(), sd[b + "norm2.bias"].float(), 1e-6), em)
        h = _q(F.silu(lin(sd, b + "mlp.fc1", nx, em)), em)
        x = x + lin(sd, b + "mlp.fc2", h, em) * g_mlp
    return x


def byt5_mapper(sd, x, em):
    """text_encoder/byT5/__init__.py:207-250 ByT5Mapper(use_residual=False): LN(eps 1e-5) fc1 GELU(erf) fc2 GELU fc3."""
    p = "byt5_in."
    h = _q(F.layer_norm(x, (x.shape[-1],), sd[p + "layernorm.weight"].float(), sd[p + "layernorm.bias"].float(), 1e-5), em)
    h = _q(F.gelu(lin(sd, p + "fc1", h, em)), em)
    h = _q(F.gelu(lin(sd, p + "fc2", h, em)), em)
    return lin(sd, p + "fc3", h, em)


def rms_head(x, w, eps=1e-6):                                  # norm_layers.py:62-70 RMSNorm.apply_ over head_dim
    return x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps) * w.float()


def double_block(sd, cfg, i, img, txt, vec, cos, sin, n_txt_valid, ref_casts, em):
    """MMDoubleStreamBlock.forward (models.py:158-318). img [L,D], txt [Lt,D] (valid tokens first), vec [D]."""
    D, H = cfg["hidden_size"], cfg["heads_num"]
    hd = D // H
    rc = ref_casts or em
    out = {}
    mods = {}
    for st in ("img", "txt"):
        p = f"double_blocks.{i}.{st}_"
        mods[st] = (F.silu(vec) @ sd[p + "mod.linear.weight"].float().t() + sd[p + "mod.linear.bias"].float()).chunk(6)
    qs, ks, vs = [], [], []
    for st, x in (("img", img), ("txt", txt)):
        p = f"double_blocks.{i}.{st}_"
        sh1, sc1 = mods[st][0], mods[st][1]
        xm = F.layer_norm(x, (D,), eps=1e-6)
        if st == "img":
            xm = _q(xm, rc)                                     # img_modulated.to(torch.bfloat16), models.py:211
        xm = _q(xm * (1 + sc1) + sh1, em or (rc and st == "img"))   # modulate_ writes into the bf16 tensor (:216)
        if p + "attn_qkv.weight" in sd:                          # HunyuanVideo 1.0 keeps q|k|v fused (split by mmgp at load)
            q, k, v = (u.reshape(-1, H, hd) for u in _q(lin(sd, p + "attn_qkv", xm, em), em).chunk(3, -1))
        else:
            q = _q(lin(sd, p + "attn_q", xm, em), em).reshape(-1, H, hd)
            k = _q(lin(sd, p + "attn_k", xm, em), em).reshape(-1, H, hd)
            v = _q(lin(sd, p + "attn_v", xm, em), em).reshape(-1, H, hd)
        q, k = rms_head(q, sd[p + "attn_q_norm.weight"]), rms_head(k, sd[p + "attn_k_norm.weight"])
        if st == "img":
            q, k = apply_rope(q, cos, sin), apply_rope(k, cos, sin)
        qs.append(_q(q, em)), ks.append(_q(k, em)), vs.append(v)
    L = img.shape[0]
    n = L + n_txt_valid                                         # seqlens_q = seqlens_kv = text_len + img_len (models.py:1086-1088)
    q, k, v = (torch.cat(u, 0)[:n] for u in (qs, ks, vs))
    attn = _q(attention(q, k, v, em).reshape(n, D), em)
    attn = torch.cat([attn, attn.new_zeros(L + txt.shape[0] - n, D)], 0)
    for st, x, a in (("img", img, attn[:L]), ("txt", txt, attn[L:])):
        p = f"double_blocks.{i}.{st}_"
        sh2, sc2, g1, g2 = mods[st][3], mods[st][4], mods[st][2], mods[st][5]
        x = x + lin(sd, p + "attn_proj", a, em) * g1
        xm = _q(F.layer_norm(x, (D,), eps=1e-6), rc)            # .to(torch.bfloat16), models.py:290 / :313
        xm = _q(xm * (1 + sc2) + sh2, rc)                       # in-place modulate_ on the bf16 tensor (:291 / :314)
        h = _q(F.gelu(lin(sd, p + "mlp.fc1", xm, em), approximate="tanh"), em)
        y = lin(sd, p + "mlp.fc2", h, em)
        if st == "img":
            y = _q(y, rc)                                       # img_mlp.apply_ stores the MLP output into the bf16 buffer (:292)
        x = x + y * g2
        out[st] = x
    return out["img"], out["txt"]


def single_block(sd, cfg, i, img, txt, vec, cos, sin, n_txt_valid, ref_casts, em):
    """MMSingleStreamBlock.forward (models.py:393-508): one modulation for both streams, linear1 = q|k|v|mlp_in, joint
    attention, linear2(cat(attn, gelu(mlp))) stored into the bf16 buffer, gated accumulate."""
    D, H = cfg["hidden_size"], cfg["heads_num"]
    hd = D // H
    rc = ref_casts or em
    p = f"single_blocks.{i}."
    sh, sc, gate = (F.silu(vec) @ sd[p + "modulation.linear.weight"].float().t() + sd[p + "modulation.linear.bias"].float()).chunk(3)
    L = img.shape[0]
    x = torch.cat([img, txt], 0)
    xm = _q(_q(F.layer_norm(x, (D,), eps=1e-6), rc) * (1 + sc) + sh, rc)      # pre_norm -> bf16 -> modulate_ (:427-435)
    h1 = lin(sd, p + "linear1", xm, em)
    q, k, v = (_q(u, em).reshape(-1, H, hd) for u in h1[:, :3 * D].chunk(3, -1))
    q, k = rms_head(q, sd[p + "q_norm.weight"]), rms_head(k, sd[p + "k_norm.weight"])
    q = torch.cat([apply_rope(q[:L], cos, sin), q[L:]], 0)
    k = torch.cat([apply_rope(k[:L], cos, sin), k[L:]], 0)
    n = L + n_txt_valid
    attn = _q(attention(_q(q, em)[:n], _q(k, em)[:n], v[:n], em).reshape(n, D), em)
    attn = torch.cat([attn, attn.new_zeros(x.shape[0] - n, D)], 0)
    mlp = _q(F.gelu(h1[:, 3 * D:], approximate="tanh"), em)
    y = _q(lin(sd, p + "linear2", torch.cat([attn, mlp], -1), em), rc)       # x_chunk[...] = linear2(...) into bf16 (:493)
    x = x + y * gate
    return x[:L], x[L:]


def vision_projection(sd, vs, em):
    """VisionProjection (embed_layers.py:62-77): LayerNorm -> Linear -> GELU -> Linear -> LayerNorm (eps 1e-5, affine)."""
    Dv, D = vs.shape[-1], sd["vision_in.proj.3.weight"].shape[0]
    h = _q(F.layer_norm(vs, (Dv,), sd["vision_in.proj.0.weight"].float(), sd["vision_in.proj.0.bias"].float(), 1e-5), em)
    h = _q(F.gelu(lin(sd, "vision_in.proj.1", h, em)), em)
    h = lin(sd, "vision_in.proj.3", h, em)
    return _q(F.layer_norm(h, (D,), sd["vision_in.proj.4.weight"].float(), sd["vision_in.proj.4.bias"].float(), 1e-5), em)


def hy_forward(sd, cfg, x, t, text_states, text_mask, byt5_states=None, byt5_mask=None, freqs=None, ref_casts=True,
               emulate_bf16=False, num_blocks=None, text_states_2=None, guidance=None, vision_states=None):
    """x [1,Cin,T,H,W] -> [1,Cout,T,H,W]; masks must mark a valid PREFIX (as the reference's encoders produce)."""
    em = emulate_bf16
    D = cfg["hidden_size"]
    _, Cin, T, H, W = x.shape
    P = cfg["patch_size"][1]
    cos, sin = freqs if freqs is not None else rope_tables_hy((T, H // P, W // P))
    vec = timestep_embedder(sd, "time_in.", t)[0]
    if text_states_2 is not None:                              # pooled-text vector (models.py:1012-1019, MLPEmbedder)
        h = F.silu(text_states_2[0].float() @ sd["vector_in.in_layer.weight"].float().t() + sd["vector_in.in_layer.bias"].float())
        vec = vec + h @ sd["vector_in.out_layer.weight"].float().t() + sd["vector_in.out_layer.bias"].float()
    if guidance is not None:                                   # guidance-distilled models (models.py:1021-1029)
        vec = vec + timestep_embedder(sd, "guidance_in.", guidance)[0]
    # PatchEmbed (embed_layers.py:9-60): Conv3d k = s = patch; K order (c, ph, pw), tokens (t, h', w')
    xp = x[0].reshape(Cin, T, H // P, P, W // P, P).permute(1, 2, 4, 0, 3, 5).reshape(T * (H // P) * (W // P), Cin * P * P)
    img = xp @ sd["img_in.proj.weight"].float().reshape(D, -1).t() + sd["img_in.proj.bias"].float()
    nt = int(text_mask[0].sum())
    txt = token_refiner(sd, cfg, text_states[0, :nt].float(), t, nt, em)
    n_valid, n_pad = nt, text_states.shape[1] - nt
    if "cond_type_embedding.weight" in sd:
        txt = txt + sd["cond_type_embedding.weight"][0].float()
    if byt5_states is not None:
        nb = int(byt5_mask[0].sum())
        b5 = byt5_mapper(sd, byt5_states[0, :nb].float(), em) + sd["cond_type_embedding.weight"][1].float()
        txt = torch.cat([b5, txt], 0)                           # reorder_txt_token(zero_feat=True), models.py:910-935
        n_valid, n_pad = n_valid + nb, n_pad + byt5_states.shape[1] - nb
    if vision_states is not None:                              # models.py:1063-1071: projected image-encoder tokens in front, cond type 2
        vis = vision_projection(sd, vision_states[0].float(), em) + sd["cond_type_embedding.weight"][2].float()
        txt = torch.cat([vis, txt], 0)
        n_valid += vis.shape[0]
    txt = torch.cat([txt, txt.new_zeros(n_pad, D)], 0)
    nblk = cfg["mm_double_blocks_depth"] if num_blocks is None else num_blocks
    for i in range(nblk):
        img, txt = double_block(sd, cfg, i, img, txt, vec, cos, sin, n_valid, ref_casts, em)
    for i in range(cfg.get("mm_single_blocks_depth", 0)):
        img, txt = single_block(sd, cfg, i, img, txt, vec, cos, sin, n_valid, ref_casts, em)
    sh, sc = (F.silu(vec) @ sd["final_layer.adaLN_modulation.1.weight"].float().t() + sd["final_layer.adaLN_modulation.1.bias"].float()).chunk(2)
    y = _q(F.layer_norm(img, (D,), eps=1e-6) * (1 + sc) + sh, em)
    y = lin(sd, "final_layer.linear", y, em)                     # [L, Cout * P * P], feature order (c, ph, pw)
    C = cfg["out_channels"]
    y = y.reshape(T, H // P, W // P, C, 1, P, P)
    return torch.einsum("thwcopq->ctohpwq", y).reshape(1, C, T, H, W)   # unpatchify, models.py:1235-1248
