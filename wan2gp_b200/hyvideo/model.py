"""B200-native HYVideoDiffusionTransformer (hot-path rows H1-H4 of SURVEY.md section 8a) for the 'HYVideo-1_5' family (54
double-stream blocks) and the HunyuanVideo 1.0 family ('HYVideo-T/2[-cfgdistill]': 20 double + 40 single-stream blocks,
fused qkv, pooled-text and guidance vectors, patch (1,2,2)):
same constructor arguments, state-dict names and forward() contract as the reference
`models/hyvideo/modules/models.py::HYVideoDiffusionTransformer` (:946-1233) -- double-stream blocks (:158-318) with per-head
QK RMSNorm, RoPE on the image stream, joint attention over cat(img, txt) trimmed to the valid text length, token refiner
(token_refiner.py:165-237), byT5 mapper, cond-type embedding, adaLN final layer and unpatchify.

All arithmetic runs in the same sm_100a kernels as the Wan path (tcgen05 GEMM with fused bias/activation/gate/residual,
TMEM flash attention, row kernels); the residual streams are fp32, GEMM operands bf16.
Joint attention needs no concatenation kernels: the img and txt q|k|v GEMMs write into one [L + Lt, 3D] buffer and the
attention kernel reads its first L + n_valid rows.
"""
import torch

from .. import ops

bf16, f32 = torch.bfloat16, torch.float32

ACT_GELU_TANH, ACT_SILU, ACT_GELU_ERF = 1, 2, 3


def get_rotary_pos_embed(latents_size, rope_dim_list=(16, 56, 56), theta=256.0, enable_riflex=False, k=4, L_test=66):
    """hunyuan.py:677-725 -> hyvideo/modules/posemb_layers.py::get_nd_rotary_pos_embed(theta=256, use_real=True):
    (cos, sin) fp32 [T*H*W, 128] over the patch grid `latents_size` = (T, H, W) after patching."""
    import math
    grids = torch.meshgrid(*[torch.arange(int(n), dtype=torch.float32) for n in latents_size], indexing="ij")
    cos, sin = [], []
    for axis, (d, g) in enumerate(zip(rope_dim_list, grids)):
        freqs = 1.0 / (theta ** (torch.arange(0, d, 2, dtype=torch.float32)[: d // 2] / d))
        if axis == 0 and enable_riflex:
            freqs[k - 1] = 0.9 * 2 * math.pi / L_test
        ang = torch.outer(g.reshape(-1), freqs)
        cos.append(ang.cos().repeat_interleave(2, dim=1))
        sin.append(ang.sin().repeat_interleave(2, dim=1))
    return torch.cat(cos, 1), torch.cat(sin, 1)


class _Stream:
    __slots__ = ("mod_w", "mod_b", "w_qkv", "b_qkv", "qn", "kn", "w_proj", "b_proj", "w_fc1", "b_fc1", "w_fc2", "b_fc2")


class _Single:
    __slots__ = ("mod_w", "mod_b", "w1", "b1", "w2", "b2", "qn", "kn")


class HYVideoDiffusionTransformer(torch.nn.Module):
    def __init__(self, i2v_condition_type=None, patch_size=(1, 2, 2), in_channels=4, out_channels=None, hidden_size=3072,
                 heads_num=24, mlp_width_ratio=4.0, mlp_act_type="gelu_tanh", mm_double_blocks_depth=20,
                 mm_single_blocks_depth=40, rope_dim_list=(16, 56, 56), qkv_bias=True, qk_norm=True, qk_norm_type="rms",
                 guidance_embed=False, text_projection="single_refiner", use_attention_mask=True, text_states_dim=4096,
                 text_states_dim_2=768, text_pool_type=True, glyph_byT5_v2=False, use_cond_type_embedding=False,
                 use_meanflow=False, vision_projection=False, vision_states_dim=1280, pre_split_qkv=False, device="cuda", **unused):
        super().__init__()
        if use_meanflow or i2v_condition_type not in (None, "latent_concat"):   # "latent_concat" acts outside the model (extra input channels)
            raise NotImplementedError("meanflow / token-replace (i2v) conditioning is outside the hot path")
        if vision_projection not in (False, None, "linear"):
            raise NotImplementedError(f"vision_projection {vision_projection!r}")
        if hidden_size // heads_num != 128 or not qk_norm or qk_norm_type != "rms" or text_projection != "single_refiner":
            raise NotImplementedError("HY hot path requires head_dim 128, RMS qk-norm and the single_refiner text projection")
        if tuple(patch_size) not in ((1, 1, 1), (1, 2, 2)):
            raise NotImplementedError(f"patch_size {patch_size}")
        self.patch_size, self.in_channels = list(patch_size), in_channels
        self.out_channels = in_channels if out_channels is None else out_channels
        self.hidden_size, self.heads_num, self.depth = hidden_size, heads_num, mm_double_blocks_depth
        self.single_depth, self.mlp_hidden = mm_single_blocks_depth, int(hidden_size * mlp_width_ratio)
        self.text_pool_type, self.text_states_dim_2 = text_pool_type, text_states_dim_2
        self.rope_dim_list, self.text_states_dim = list(rope_dim_list), text_states_dim
        self.glyph_byT5_v2, self.use_cond = glyph_byT5_v2, use_cond_type_embedding
        self.i2v_condition_type, self.guidance_embed = i2v_condition_type, guidance_embed
        self.vision_projection, self.vision_states_dim = vision_projection or None, vision_states_dim
        self.device = torch.device(device)
        self.cache = None
        self.double_blocks, self.single_blocks = [], []
        self._g, self._ref = {}, []
        self._ready = False

    # ------------------------------------------------------------------ weights
    def _d(self, t, dtype):
        return t.detach().to(self.device, dtype).contiguous()

    def _lin(self, sd, name, dtype=bf16):
        return self._d(sd[name + ".weight"], dtype), self._d(sd[name + ".bias"], f32)

    def _pack_stream(self, sd, p):
        s = _Stream()
        s.mod_w, s.mod_b = self._lin(sd, p + "mod.linear", f32)
        if p + "attn_qkv.weight" in sd:        # HunyuanVideo 1.0 checkpoints keep q|k|v fused (hunyuan_handler.py:274-278)
            s.w_qkv, s.b_qkv = self._lin(sd, p + "attn_qkv")
        else:
            s.w_qkv = self._d(torch.cat([sd[p + f"attn_{l}.weight"] for l in "qkv"], 0), bf16)
            s.b_qkv = self._d(torch.cat([sd[p + f"attn_{l}.bias"] for l in "qkv"], 0), f32)
        s.qn, s.kn = self._d(sd[p + "attn_q_norm.weight"], f32), self._d(sd[p + "attn_k_norm.weight"], f32)
        s.w_proj, s.b_proj = self._lin(sd, p + "attn_proj")
        s.w_fc1, s.b_fc1 = self._lin(sd, p + "mlp.fc1")
        s.w_fc2, s.b_fc2 = self._lin(sd, p + "mlp.fc2")
        return s

    def load_state_dict(self, sd, strict=True, assign=False):
        g, D = self._g, self.hidden_size
        g["img_w"] = self._d(sd["img_in.proj.weight"].reshape(D, -1), f32)
        g["img_b"] = self._d(sd["img_in.proj.bias"], f32)
        for n in ("time_in.mlp.0", "time_in.mlp.2", "txt_in.t_embedder.mlp.0", "txt_in.t_embedder.mlp.2",
                  "txt_in.c_embedder.linear_1", "txt_in.c_embedder.linear_2", "final_layer.adaLN_modulation.1"):
            g[n] = self._lin(sd, n, f32)
        g["txt_embed"] = self._lin(sd, "txt_in.input_embedder")
        g["final"] = self._lin(sd, "final_layer.linear")
        self._ref = []
        for j in range(2):
            p = f"txt_in.individual_token_refiner.blocks.{j}."
            self._ref.append({"ada": self._lin(sd, p + "adaLN_modulation.1", f32),
                              "n1": (self._d(sd[p + "norm1.weight"], f32), self._d(sd[p + "norm1.bias"], f32)),
                              "n2": (self._d(sd[p + "norm2.weight"], f32), self._d(sd[p + "norm2.bias"], f32)),
                              "qkv": self._lin(sd, p + "self_attn_qkv"), "proj": self._lin(sd, p + "self_attn_proj"),
                              "fc1": self._lin(sd, p + "mlp.fc1"), "fc2": self._lin(sd, p + "mlp.fc2")})
        if self.glyph_byT5_v2:
            g["byt5_ln"] = (self._d(sd["byt5_in.layernorm.weight"], f32), self._d(sd["byt5_in.layernorm.bias"], f32))
            for n in ("fc1", "fc2", "fc3"):
                g["byt5_" + n] = self._lin(sd, "byt5_in." + n)
        if self.use_cond:
            g["cond"] = self._d(sd["cond_type_embedding.weight"], f32)
        g.pop("vision", None)
        if self.vision_projection == "linear" and "vision_in.proj.1.weight" in sd:       # t2v checkpoints may omit it (only used with vision_states)
            g["vision"] = {"ln0": (self._d(sd["vision_in.proj.0.weight"], f32), self._d(sd["vision_in.proj.0.bias"], f32)),
                           "fc1": self._lin(sd, "vision_in.proj.1"), "fc2": self._lin(sd, "vision_in.proj.3"),
                           "ln1": (self._d(sd["vision_in.proj.4.weight"], f32), self._d(sd["vision_in.proj.4.bias"], f32))}
        if self.text_pool_type is not None:
            g["vector_in.in_layer"] = self._lin(sd, "vector_in.in_layer", f32)
            g["vector_in.out_layer"] = self._lin(sd, "vector_in.out_layer", f32)
        if self.guidance_embed:
            g["guidance_in.mlp.0"] = self._lin(sd, "guidance_in.mlp.0", f32)
            g["guidance_in.mlp.2"] = self._lin(sd, "guidance_in.mlp.2", f32)
        self.double_blocks = [(self._pack_stream(sd, f"double_blocks.{i}.img_"), self._pack_stream(sd, f"double_blocks.{i}.txt_"))
                              for i in range(self.depth)]
        self.single_blocks = []
        for i in range(self.single_depth):
            p, b = f"single_blocks.{i}.", _Single()
            b.mod_w, b.mod_b = self._lin(sd, p + "modulation.linear", f32)
            b.w1, b.b1 = self._lin(sd, p + "linear1")          # rows: q | k | v | mlp_in  (split by mmgp in the reference)
            b.w2, b.b2 = self._lin(sd, p + "linear2")
            b.qn, b.kn = self._d(sd[p + "q_norm.weight"], f32), self._d(sd[p + "k_norm.weight"], f32)
            self.single_blocks.append(b)
        self._ready = True
        return torch.nn.modules.module._IncompatibleKeys([], [])

    def init_synthetic(self, seed=0):
        from .. import synth
        cfg = dict(hidden_size=self.hidden_size, heads_num=self.heads_num, mlp_width_ratio=4, mm_double_blocks_depth=self.depth,
                   in_channels=self.in_channels, out_channels=self.out_channels, text_states_dim=self.text_states_dim,
                   patch_size=self.patch_size, mm_single_blocks_depth=self.single_depth, guidance_embed=self.guidance_embed,
                   text_states_dim_2=self.text_states_dim_2, family="1.0" if self.text_pool_type is not None else "1.5")
        self.load_state_dict({n: synth.make_hy_tensor(n, s, seed, self.device) for n, s in synth.hy_param_shapes(cfg).items()})
        return self

    # ------------------------------------------------------------------ pieces
    def _tembed(self, key, t):
        """TimestepEmbedder (embed_layers.py:137-174): sinusoid(256) -> Linear -> SiLU -> Linear, fp32."""
        (w0, b0), (w2, b2) = self._g[key + ".mlp.0"], self._g[key + ".mlp.2"]
        s = ops.sinusoid(float(t), 256, self.device)
        return ops.gemv(ops.gemv(s, w0, b0, silu_out=True), w2, b2)

    def _refiner(self, txt, t):
        """SingleTokenRefiner on the VALID tokens txt [n, text_dim] fp32 -> [n, D] fp32 (token_refiner.py:165-237)."""
        g, D, H = self._g, self.hidden_size, self.heads_num
        ta = self._tembed("txt_in.t_embedder", t)
        (w1, b1), (w2, b2) = g["txt_in.c_embedder.linear_1"], g["txt_in.c_embedder.linear_2"]
        c = ops.add_vec(ta, ops.gemv(ops.gemv(ops.col_mean(txt), w1, b1, silu_out=True), w2, b2))
        x = ops.gemm(ops.cast_bf16(txt), g["txt_embed"][0], bias=g["txt_embed"][1], out_dtype=f32)
        for r in self._ref:
            mod = ops.gemv(c, r["ada"][0], r["ada"][1], silu_in=True)
            nx = ops.ln_modulate(x, r["n1"][1], r["n1"][0], affine=True)
            qkv = ops.gemm(nx, r["qkv"][0], bias=r["qkv"][1])
            a = ops.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], H)
            ops.gemm(a, r["proj"][0], out=x, bias=r["proj"][1], gate=mod[:D], accumulate=True)
            nx = ops.ln_modulate(x, r["n2"][1], r["n2"][0], affine=True)
            h = ops.gemm(nx, r["fc1"][0], bias=r["fc1"][1], act=ACT_SILU)
            ops.gemm(h, r["fc2"][0], out=x, bias=r["fc2"][1], gate=mod[D:], accumulate=True)
        return x

    def _byt5(self, b5):
        """ByT5Mapper(use_residual=False) (text_encoder/byT5/__init__.py:207-250): LN(1e-5) fc1 GELU fc2 GELU fc3."""
        g = self._g
        h = ops.ln_modulate(b5, g["byt5_ln"][1], g["byt5_ln"][0], affine=True, eps=1e-5)
        h = ops.gemm(h, g["byt5_fc1"][0], bias=g["byt5_fc1"][1], act=ACT_GELU_ERF)
        h = ops.gemm(h, g["byt5_fc2"][0], bias=g["byt5_fc2"][1], act=ACT_GELU_ERF)
        return ops.gemm(h, g["byt5_fc3"][0], bias=g["byt5_fc3"][1], out_dtype=f32)

    def _vision(self, vs):
        """VisionProjection (embed_layers.py:62-77): LayerNorm -> Linear -> GELU(erf) -> Linear -> LayerNorm, vs fp32 [n, vision_dim] -> fp32 [n, D]."""
        v = self._g["vision"]
        h = ops.ln_modulate(vs, v["ln0"][1], v["ln0"][0], affine=True, eps=1e-5)
        h = ops.gemm(h, v["fc1"][0], bias=v["fc1"][1], act=ACT_GELU_ERF)
        h = ops.gemm(h, v["fc2"][0], bias=v["fc2"][1], out_dtype=f32)
        return ops.ln_modulate(h, v["ln1"][1], v["ln1"][0], affine=True, eps=1e-5).float()       # bf16 tokens, as the reference's (bf16) stream holds them

    def _text(self, text_states, text_mask, byt5_states, byt5_mask, t, vision_states=None):
        """txt_in + cond-type embedding + byT5 + reorder_txt_token(zero_feat=True) (+ the projected image-encoder tokens in front,
        models.py:1036-1071, 910-935): returns the fp32 text stream [Lt, D] ordered [vision | byT5 valid | LLM valid | zero padding] and the
        valid length."""
        D = self.hidden_size
        tm = text_mask.bool().cpu() if text_mask is not None else torch.ones(text_states.shape[0], dtype=torch.bool)
        valid = self._refiner(text_states.to(self.device, f32)[tm.to(self.device)].contiguous(), t)
        if self.use_cond:
            valid = ops.add_vec(valid.reshape(-1), self._g["cond"][0]).reshape(-1, D)
        parts, n_total = [valid], text_states.shape[0]
        if self.glyph_byT5_v2 and byt5_states is not None:
            bm = byt5_mask.bool().cpu()
            n_total += byt5_states.shape[0]
            if int(bm.sum()) > 0:
                b = self._byt5(byt5_states.to(self.device, f32)[bm.to(self.device)].contiguous())
                if self.use_cond:
                    b = ops.add_vec(b.reshape(-1), self._g["cond"][1]).reshape(-1, D)
                parts.insert(0, b)
        if vision_states is not None:                        # models.py:1063-1071: all vision tokens are valid, cond type 2, in front
            v = self._vision(vision_states.to(self.device, f32).contiguous())
            if self.use_cond:
                v = ops.add_vec(v.reshape(-1), self._g["cond"][2]).reshape(-1, D)
            n_total += v.shape[0]
            parts.insert(0, v)
        n_valid = sum(p.shape[0] for p in parts)
        txt = torch.zeros(n_total, D, device=self.device, dtype=f32)
        txt[:n_valid] = torch.cat(parts, 0)
        return txt, n_valid

    def _double_block(self, blk, img, txt, vec, cos, sin, n_valid, qkv, attn):
        D, H = self.hidden_size, self.heads_num
        L = img.shape[0]
        mods = []
        for s, x, rows in ((blk[0], img, slice(0, L)), (blk[1], txt, slice(L, None))):
            m = ops.gemv(vec, s.mod_w, s.mod_b, silu_in=True)              # ModulateDiT (modulate_layers.py:27-33)
            mods.append(m)
            a = ops.ln_modulate(x, m[0:D], m[D:2 * D], pre_round=True)     # norm1 -> bf16 -> modulate_ (models.py:210-216)
            ops.gemm(a, s.w_qkv, out=qkv[rows], bias=s.b_qkv)
            rope = (cos, sin) if s is blk[0] else (None, None)             # RoPE on the image stream only (:232-235)
            ops.qk_rmsnorm_rope_(qkv[rows, :D], qkv[rows, D:2 * D], s.qn, s.kn, 1e-6, *rope, per_head=True)
        n = L + n_valid                                                    # q_lens = k_lens = img_len + text_len (:1086-1088)
        ops.attention(qkv[:n, :D], qkv[:n, D:2 * D], qkv[:n, 2 * D:], H, out=attn[:n])
        for s, x, rows, m in ((blk[0], img, slice(0, L), mods[0]), (blk[1], txt, slice(L, None), mods[1])):
            ops.gemm(attn[rows], s.w_proj, out=x, bias=s.b_proj, gate=m[2 * D:3 * D], accumulate=True)
            a = ops.ln_modulate(x, m[3 * D:4 * D], m[4 * D:5 * D], pre_round=True)
            h = ops.gemm(a, s.w_fc1, bias=s.b_fc1, act=ACT_GELU_TANH)
            ops.gemm(h, s.w_fc2, out=x, bias=s.b_fc2, gate=m[5 * D:], accumulate=True)

    def _single_block(self, b, img, txt, vec, cos, sin, n_valid, qkv, cat):
        """MMSingleStreamBlock (models.py:393-508): shared modulation, linear1 = q|k|v|mlp_in, joint attention written straight
        into the first D columns of the [attn | gelu(mlp)] buffer that feeds linear2, gated accumulate into both streams."""
        D, H, L = self.hidden_size, self.heads_num, img.shape[0]
        m = ops.gemv(vec, b.mod_w, b.mod_b, silu_in=True)
        for x, rows, rope in ((img, slice(0, L), (cos, sin)), (txt, slice(L, None), (None, None))):
            a = ops.ln_modulate(x, m[0:D], m[D:2 * D], pre_round=True)
            ops.gemm(a, b.w1[:3 * D], out=qkv[rows], bias=b.b1[:3 * D])
            ops.gemm(a, b.w1[3 * D:], out=cat[rows, D:], bias=b.b1[3 * D:], act=ACT_GELU_TANH)
            ops.qk_rmsnorm_rope_(qkv[rows, :D], qkv[rows, D:2 * D], b.qn, b.kn, 1e-6, *rope, per_head=True)
        n = L + n_valid
        ops.attention(qkv[:n, :D], qkv[:n, D:2 * D], qkv[:n, 2 * D:], H, out=cat[:n, :D])
        for x, rows in ((img, slice(0, L)), (txt, slice(L, None))):
            ops.gemm(cat[rows], b.w2, out=x, bias=b.b2, gate=m[2 * D:], accumulate=True)

    # ------------------------------------------------------------------ forward (reference contract)
    @torch.no_grad()
    def forward(self, x, t, ref_latents=None, text_states=None, text_mask=None, text_states_2=None, freqs_cos=None,
                freqs_sin=None, guidance=None, pipeline=None, x_id=0, step_no=0, callback=None, audio_prompts=None,
                motion_exp=None, motion_pose=None, fps=None, face_mask=None, audio_strength=None, bg_latents=None,
                vision_states=None, byt5_text_states=None, byt5_text_mask=None, timesteps_r=None):
        """x [B,Cin,T,H,W], t [B], text_states [B,Lt,text_dim], text_mask [B,Lt], byt5_text_states [B,Lb,1472] ->
        [B,Cout,T,H,W] fp32, or None when `pipeline._interrupt` is raised (polled once per block, models.py:1146-1149)."""
        if not self._ready:
            raise RuntimeError("HYVideoDiffusionTransformer: load_state_dict() / init_synthetic() must be called before forward()")
        if (text_states_2 is None) != (self.text_pool_type is None):
            raise ValueError("text_states_2 must be given exactly when the model has a pooled-text vector_in (text_pool_type)")
        if self.guidance_embed and guidance is None:
            raise ValueError("Didn't get guidance strength for guidance distilled model.")       # models.py:1023-1026
        for name, v in (("ref_latents", ref_latents), ("audio_prompts", audio_prompts),
                        ("motion_exp", motion_exp), ("motion_pose", motion_pose), ("fps", fps), ("bg_latents", bg_latents),
                        ("timesteps_r", timesteps_r)):
            if v is not None:
                raise NotImplementedError(f"HYVideoDiffusionTransformer.forward: `{name}` is outside the hot path")
        if vision_states is not None and "vision" not in self._g:
            raise ValueError("vision_states given but the model has no vision_in projection (vision_projection='linear' + its weights)")
        B, Cin, T, H, W = x.shape
        P, D = self.patch_size[1], self.hidden_size
        L = T * (H // P) * (W // P)
        if freqs_cos is None:
            freqs_cos, freqs_sin = get_rotary_pos_embed((T, H // P, W // P), self.rope_dim_list)
        cos, sin = freqs_cos.to(self.device, f32).contiguous(), freqs_sin.to(self.device, f32).contiguous()
        outs = []
        streams = []
        for i in range(B):
            ti = float(t.flatten()[i] if t.numel() > 1 else t.flatten()[0])
            vec = self._tembed("time_in", ti)
            if text_states_2 is not None:                       # vector_in MLPEmbedder (models.py:1012-1019)
                (w1, b1), (w2, b2) = self._g["vector_in.in_layer"], self._g["vector_in.out_layer"]
                pooled = text_states_2[i].to(self.device, f32).contiguous()
                vec = ops.add_vec(vec, ops.gemv(ops.gemv(pooled, w1, b1, silu_out=True), w2, b2))
            if self.guidance_embed:                             # guidance_in TimestepEmbedder (models.py:1021-1029)
                gi = float(guidance.flatten()[i] if guidance.numel() > 1 else guidance.flatten()[0])
                vec = ops.add_vec(vec, self._tembed("guidance_in", gi))
            img = ops.patch_embed(x[i].to(self.device, f32).contiguous(), None, self._g["img_w"], self._g["img_b"], D, patch=P)
            txt, n_valid = self._text(text_states[i], None if text_mask is None else text_mask[i],
                                      None if byt5_text_states is None else byt5_text_states[i],
                                      None if byt5_text_mask is None else byt5_text_mask[i], ti,
                                      None if vision_states is None else vision_states[0])        # the reference repeats ONE image's tokens over the batch (:1064)
            Lt = txt.shape[0]
            cat = torch.zeros(L + Lt, D + self.mlp_hidden, device=self.device, dtype=bf16) if self.single_blocks else None
            streams.append((img, txt, vec, n_valid, torch.empty(L + Lt, 3 * D, device=self.device, dtype=bf16),
                            torch.zeros(L + Lt, D, device=self.device, dtype=bf16), cat))
        for blk in self.double_blocks:
            for (img, txt, vec, n_valid, qkv, attn, cat) in streams:
                if callback is not None:
                    callback(-1, None, False, True)
                if pipeline is not None and getattr(pipeline, "_interrupt", False):
                    return None
                self._double_block(blk, img, txt, vec, cos, sin, n_valid, qkv, attn)
        for blk in self.single_blocks:
            for (img, txt, vec, n_valid, qkv, attn, cat) in streams:
                if callback is not None:
                    callback(-1, None, False, True)
                if pipeline is not None and getattr(pipeline, "_interrupt", False):
                    return None
                self._single_block(blk, img, txt, vec, cos, sin, n_valid, qkv, cat)
        fw, fb = self._g["final_layer.adaLN_modulation.1"]
        for (img, txt, vec, n_valid, qkv, attn, cat) in streams:
            m = ops.gemv(vec, fw, fb, silu_in=True)                          # FinalLayer (mlp_layers.py:127-131): shift, scale
            y = ops.ln_modulate(img, m[:D], m[D:])
            o = ops.gemm(y, self._g["final"][0], bias=self._g["final"][1], out_dtype=f32)
            outs.append(ops.unpatchify(o, self.out_channels, T, H, W, patch=P, c_major=True))
        return torch.stack(outs, 0)
