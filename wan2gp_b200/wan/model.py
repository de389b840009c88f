"""B200-native WanModel: same constructor arguments, state-dict names and forward() contract as the
reference `models/wan/modules/model.py::WanModel` (SURVEY.md section 8b level 3) for the plain t2v / i2v2_2
path; every FLOP runs in the hand-written sm_100a kernels of libwan2gp_b200.so (C ABI, ../ops.py).

Data layout in HBM (one sample):
  residual stream x          fp32 [L, D]           (kept fp32 end to end; the reference keeps bf16 unless
                                                     mixed_precision_transformer, any2video.py:190)
  GEMM operands              bf16: a / c / f [L, D], fused qkv [L, 3D], attention out [L, D], ffn hidden [L, F]
  weights                    bf16 nn.Linear layout [out, in]; q|k|v and cross k|v concatenated along `out`
  biases / norms / modulation / RoPE tables   fp32

Per block (reference model.py:631-711) the launches are:
  add_vec (modulation + e0) -> ln_modulate -> GEMM qkv -> qk_rmsnorm_rope(q, k) -> attention ->
  GEMM o (+bias, *gate, += x) -> ln_modulate(affine) -> GEMM q' -> rmsnorm -> GEMM kv'(ctx) -> rmsnorm ->
  attention(512 keys) -> GEMM o' (+= x) -> ln_modulate -> GEMM ffn.0 (+GELU) -> GEMM ffn.2 (*gate, += x)
"""
import math

import torch

from .. import ops
from .rope import get_rotary_pos_embed

try:  # the reference publishes per-step state through mmgp.offload.shared_state (model.py:1794-1796, 1994)
    from mmgp import offload as _offload
    shared_state = _offload.shared_state
except Exception:  # mmgp is not part of this repo; keep the same dict contract locally
    shared_state = {}

bf16, f32 = torch.bfloat16, torch.float32

_UNSUPPORTED_KWARGS = (
    "vace_context", "clip_fea", "cam_emb", "audio_proj", "multitalk_audio", "multitalk_masks", "standin_ref",
    "pose_latents", "face_pixel_values", "lynx_ip_embeds", "lynx_ref_buffer", "steadydancer_condition",
    "scail_pose_latents", "scail2_ref_latents", "scail2_pose_latents", "kiwi_source_condition", "kiwi_ref_condition",
    "bernini_sources", "vista", "shotplan_cut_frames", "animate2_ref_x", "perturbation_layers")


class _Block:
    """Packed device weights of one WanAttentionBlock (model.py:506-551)."""
    __slots__ = ("modulation", "w_qkv", "b_qkv", "nq", "nk", "w_o", "b_o", "n3_w", "n3_b", "w_cq", "b_cq", "cnq",
                 "cnk", "w_ckv", "b_ckv", "w_co", "b_co", "w1", "b1", "w2", "b2")


class _PromptCache:
    """Step-invariant text projections of at most `max_prompts` prompts (one CFG pair): the text embedding (model.py:1856) and every
    block's cross-attention K/V (model.py:255-258).  A prompt is identified by (data_ptr, version, shape) of the tensor the caller
    passes; the cache keeps a reference to that tensor so its address cannot be recycled while the entry lives.  Pure Python (unit-tested
    on CPU); eviction is oldest-first."""

    def __init__(self, max_prompts=2):
        self.max_prompts, self.emb, self.ref, self.ckv = max_prompts, {}, {}, {}

    @staticmethod
    def key(t):
        return (t.data_ptr(), t._version, tuple(t.shape))

    def get_emb(self, key):
        return self.emb.get(key)

    def put_emb(self, key, emb, ref):
        while len(self.emb) >= self.max_prompts:
            old = next(iter(self.emb))
            self.emb.pop(old), self.ref.pop(old, None)
            self.ckv = {k: v for k, v in self.ckv.items() if k[0] != old}
        self.emb[key], self.ref[key] = emb, ref

    def get_ckv(self, key, idx):
        return self.ckv.get((key, idx))

    def put_ckv(self, key, idx, ckv):
        if key in self.emb:
            self.ckv[(key, idx)] = ckv

    def clear(self):
        self.emb, self.ref, self.ckv = {}, {}, {}


class WanModel(torch.nn.Module):
    """Drop-in for the reference WanModel on the t2v / i2v2_2 path."""

    def __init__(self, model_type="t2v", patch_size=(1, 2, 2), text_len=512, in_dim=16, dim=2048, ffn_dim=8192,
                 freq_dim=256, text_dim=4096, out_dim=16, num_heads=16, num_layers=32, window_size=(-1, -1),
                 qk_norm=True, cross_attn_norm=True, eps=1e-6, device="cuda", **unused):
        super().__init__()
        if model_type not in ("t2v", "i2v2_2", "ti2v2_2"):
            raise NotImplementedError(f"model_type {model_type!r}: only the t2v / i2v2_2 hot path is implemented")
        if tuple(patch_size) != (1, 2, 2) or dim % num_heads or dim // num_heads != 128 or not qk_norm or not cross_attn_norm:
            raise NotImplementedError("Wan hot path requires patch (1,2,2), head_dim 128, qk_norm and cross_attn_norm")
        self.model_type, self.patch_size, self.text_len = model_type, tuple(patch_size), text_len
        self.in_dim, self.dim, self.ffn_dim, self.freq_dim = in_dim, dim, ffn_dim, freq_dim
        self.text_dim, self.out_dim, self.num_heads, self.num_layers, self.eps = text_dim, out_dim, num_heads, num_layers, eps
        self.device = torch.device(device)
        self.cache = None                       # TeaCache/MagCache slot read by callers; step skipping is out of scope
        self._lock_dtype = torch.float32
        self.blocks = []                        # len(model.blocks) is read by callers
        self._g = {}                            # global (non-block) packed weights
        self._freqs_cache = {}
        self._prompts = _PromptCache()
        self._stacked_ckv = {}                  # (prompt keys of all streams, block) -> their text K|V stacked along the rows
        self.cache_context = False              # reuse step-invariant text projections across steps: the text
                                                # embedding (model.py:1856) and every block's cross-attention K/V (model.py:255-258)
        # optional: one CUDA graph per transformer block (launch-bound small configs); the per-block
        # interrupt poll of the reference stays between graph replays
        self.use_cuda_graphs = False
        self._graphs = {}
        self._ready = False

    # ------------------------------------------------------------------ weights
    @staticmethod
    def preprocess_key(k):
        """Checkpoint-key rewrites applied by the reference loader (model.py:914-941)."""
        if k.startswith("model.diffusion_model."):
            k = k[len("model.diffusion_model."):]
        return k.replace(".block.", ".") if k.startswith("blocks.") else k

    def _dev(self, t, dtype):
        return t.detach().to(device=self.device, dtype=dtype).contiguous()

    def _pack_globals(self, sd):
        D, g = self.dim, self._g
        g["pe_w"] = self._dev(sd["patch_embedding.weight"].reshape(D, -1), f32)
        g["pe_b"] = self._dev(sd["patch_embedding.bias"], f32)
        for i in (0, 2):
            g[f"txt_w{i}"] = self._dev(sd[f"text_embedding.{i}.weight"], bf16)
            g[f"txt_b{i}"] = self._dev(sd[f"text_embedding.{i}.bias"], f32)
            g[f"time_w{i}"] = self._dev(sd[f"time_embedding.{i}.weight"], f32)
            g[f"time_b{i}"] = self._dev(sd[f"time_embedding.{i}.bias"], f32)
        g["tproj_w"] = self._dev(sd["time_projection.1.weight"], f32)
        g["tproj_b"] = self._dev(sd["time_projection.1.bias"], f32)
        hm = sd["head.modulation.weight"] if "head.modulation.weight" in sd else sd["head.modulation"]
        g["head_mod"] = self._dev(hm.reshape(2 * D), f32)
        g["head_w"] = self._dev(sd["head.head.weight"], bf16)
        g["head_b"] = self._dev(sd["head.head.bias"], f32)

    def _pack_block(self, sd, p):
        D = self.dim
        b = _Block()
        mod = sd[p + "modulation.weight"] if (p + "modulation.weight") in sd else sd[p + "modulation"]
        b.modulation = self._dev(mod.reshape(6 * D), f32)
        sa, ca = p + "self_attn.", p + "cross_attn."
        b.w_qkv = self._dev(torch.cat([sd[sa + "q.weight"], sd[sa + "k.weight"], sd[sa + "v.weight"]], 0), bf16)
        b.b_qkv = self._dev(torch.cat([sd[sa + "q.bias"], sd[sa + "k.bias"], sd[sa + "v.bias"]], 0), f32)
        b.nq, b.nk = self._dev(sd[sa + "norm_q.weight"], f32), self._dev(sd[sa + "norm_k.weight"], f32)
        b.w_o, b.b_o = self._dev(sd[sa + "o.weight"], bf16), self._dev(sd[sa + "o.bias"], f32)
        b.n3_w, b.n3_b = self._dev(sd[p + "norm3.weight"], f32), self._dev(sd[p + "norm3.bias"], f32)
        b.w_cq, b.b_cq = self._dev(sd[ca + "q.weight"], bf16), self._dev(sd[ca + "q.bias"], f32)
        b.cnq, b.cnk = self._dev(sd[ca + "norm_q.weight"], f32), self._dev(sd[ca + "norm_k.weight"], f32)
        b.w_ckv = self._dev(torch.cat([sd[ca + "k.weight"], sd[ca + "v.weight"]], 0), bf16)
        b.b_ckv = self._dev(torch.cat([sd[ca + "k.bias"], sd[ca + "v.bias"]], 0), f32)
        b.w_co, b.b_co = self._dev(sd[ca + "o.weight"], bf16), self._dev(sd[ca + "o.bias"], f32)
        b.w1, b.b1 = self._dev(sd[p + "ffn.0.weight"], bf16), self._dev(sd[p + "ffn.0.bias"], f32)
        b.w2, b.b2 = self._dev(sd[p + "ffn.2.weight"], bf16), self._dev(sd[p + "ffn.2.bias"], f32)
        return b

    def load_state_dict(self, sd, strict=True, assign=False):
        """Accepts the reference's state-dict names in any float dtype / device and
        packs them for the kernels (QKV concatenated, bf16 GEMM operands, fp32 everything else)."""
        sd = {self.preprocess_key(k): v for k, v in sd.items() if not k.startswith("vae.")}
        self._pack_globals(sd)
        self.blocks = [self._pack_block(sd, f"blocks.{i}.") for i in range(self.num_layers)]
        self._ready = True
        self.weights_version = getattr(self, "weights_version", 0) + 1        # whole-step graphs of pipeline.WanDenoiser are keyed on it
        self._graphs = {}                                                     # captured graphs / cached projections used the old weights
        self._prompts.clear()
        self._stacked_ckv = {}
        return torch.nn.modules.module._IncompatibleKeys([], [])

    def init_synthetic(self, seed=0):
        """Random-init weights of this architecture generated ON the device, block by block (bench.py: there
        are no checkpoints offline; a 14B fp32 state dict would not fit host RAM)."""
        from .. import synth
        cfg = dict(dim=self.dim, ffn_dim=self.ffn_dim, in_dim=self.in_dim, text_dim=self.text_dim, freq_dim=self.freq_dim,
                   out_dim=self.out_dim, num_layers=self.num_layers)
        shapes = synth.wan_param_shapes(cfg)
        self._pack_globals({n: synth.make_wan_tensor(n, s, cfg, seed, self.device) for n, s in shapes.items()
                            if not n.startswith("blocks.")})
        self.blocks = []
        for i in range(self.num_layers):
            p = f"blocks.{i}."
            self.blocks.append(self._pack_block({n: synth.make_wan_tensor(n, s, cfg, seed, self.device)
                                                 for n, s in shapes.items() if n.startswith(p)}, p))
        self._ready = True
        self.weights_version = getattr(self, "weights_version", 0) + 1
        self._graphs = {}                                                     # as load_state_dict: nothing captured / cached may
        self._prompts.clear()                                                 # keep pointing at the previous weights
        self._stacked_ckv = {}
        return self

    def apply_post_init_changes(self):
        """Kept for API parity (model.py:1291-1303 wraps `modulation` parameters; nothing to do here)."""
        return self

    def lock_layers_dtypes(self, *a, **k):
        return self

    # ------------------------------------------------------------------ pieces of the forward
    def _freqs(self, freqs, thw):
        key = thw if freqs is None else (freqs[0].data_ptr(), tuple(freqs[0].shape))
        hit = self._freqs_cache.get(key)
        if hit is None:
            cos, sin = get_rotary_pos_embed(thw) if freqs is None else freqs
            hit = (cos.to(self.device, f32).contiguous(), sin.to(self.device, f32).contiguous())
            self._freqs_cache = {key: hit}
        return hit

    def _time(self, t):
        """e [D], e0 [6D]  (model.py:1815-1818)."""
        g = self._g
        if torch.is_tensor(t) and t.is_cuda:                  # whole-step graph (pipeline.WanDenoiser): the timestep stays on the device
            tval = t.reshape(-1)[:1].to(f32)
        else:
            tval = float(t.flatten()[0]) if torch.is_tensor(t) else float(t)
        s = ops.sinusoid(tval, self.freq_dim, self.device)
        h = ops.gemv(s, g["time_w0"], g["time_b0"], silu_out=True)
        e = ops.gemv(h, g["time_w2"], g["time_b2"])
        e0 = ops.gemv(e, g["tproj_w"], g["tproj_b"], silu_in=True)
        return e, e0

    def _text(self, ctx):
        """text_embedding (model.py:1133-1135, 1856): [Lt, text_dim] fp32 -> [Lt, D] bf16."""
        g = self._g
        c = ops.cast_bf16(ctx.to(self.device, f32).reshape(-1, self.text_dim))
        h = ops.gemm(c, g["txt_w0"], bias=g["txt_b0"], act=1)
        return ops.gemm(h, g["txt_w2"], bias=g["txt_b2"])

    def _block(self, b, x, e0, ctx, cos, sin, ckv=None, nseq=1):
        """One WanAttentionBlock on the fp32 residual stream x [nseq * L, D], in place (model.py:631-711).  nseq > 1: the streams of
        all CFG entries / batch items stacked along the rows -- every kernel of the block is row-wise except the two attentions, which
        take the sequence count (cos / sin then hold the tables repeated per sequence, ckv the per-sequence text K|V stacked)."""
        D, H, eps = self.dim, self.num_heads, self.eps
        m = ops.add_vec(b.modulation, e0)                                    # model.py:632
        # ---- self attention
        a = ops.ln_modulate(x, m[0:D], m[D:2 * D], eps=eps)
        qkv = ops.gemm(a, b.w_qkv, bias=b.b_qkv)
        ops.qk_rmsnorm_rope_(qkv[:, :D], qkv[:, D:2 * D], b.nq, b.nk, eps, cos, sin)
        att = ops.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], H, out=a, nseq=nseq)
        del qkv
        ops.gemm(att, b.w_o, out=x, bias=b.b_o, gate=m[2 * D:3 * D], accumulate=True)
        # ---- text cross attention
        c = ops.ln_modulate(x, b.n3_b, b.n3_w, affine=True, eps=eps, out=a)
        q = ops.gemm(c, b.w_cq, bias=b.b_cq)
        ops.rmsnorm_rope_(q, b.cnq, eps)
        if ckv is None:
            ckv = self._cross_kv(b, ctx)
        att = ops.attention(q, ckv[:, :D], ckv[:, D:], H, out=a, nseq=nseq)
        del q
        ops.gemm(att, b.w_co, out=x, bias=b.b_co, accumulate=True)
        # ---- FFN
        f = ops.ln_modulate(x, m[3 * D:4 * D], m[4 * D:5 * D], eps=eps, out=a)
        h = ops.gemm(f, b.w1, bias=b.b1, act=1)
        ops.gemm(h, b.w2, out=x, bias=b.b2, gate=m[5 * D:], accumulate=True)
        return x

    def _cross_kv(self, b, ctx):
        kv = ops.gemm(ctx, b.w_ckv, bias=b.b_ckv)
        ops.rmsnorm_rope_(kv[:, :self.dim], b.cnk, self.eps)
        return kv

    def _head(self, x, e, thw):
        """Head + unpatchify (model.py:847-865, 2100-2126)."""
        D, g = self.dim, self._g
        T, H, W = thw
        hm = ops.add_vec(g["head_mod"], e)
        y = ops.ln_modulate(x, hm[0:D], hm[D:], eps=self.eps)
        o = ops.gemm(y, g["head_w"], bias=g["head_b"], out_dtype=f32)
        return ops.unpatchify(o, self.out_dim, T, H, W)

    def _block_graphs(self, X, nseq, e0, ckv_fn, cos, sin):
        """Static input buffers + one captured CUDA graph per block for this (rows, sequences, text length) signature.
        Capture records the same C-ABI launches as the eager path (tensor maps are encoded at capture time on static
        addresses); all graphs share one memory pool, so the temporaries of one block are reused by the next.  The stacked text K|V
        of every block is a static buffer refreshed from `ckv_fn(idx)` (a cache hit when prompts are fixed)."""
        ckv0 = ckv_fn(0)
        key = (tuple(X.shape), nseq, tuple(ckv0.shape), cos.data_ptr())
        g = self._graphs.get(key)
        if g is None:
            g = {"x": torch.empty_like(X), "e0": torch.empty_like(e0), "ckv": [torch.empty_like(ckv0) for _ in self.blocks], "g": []}
            pool = torch.cuda.graph_pool_handle()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for idx, blk in enumerate(self.blocks):
                    cg = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(cg, pool=pool, stream=side):
                        self._block(blk, g["x"], g["e0"], None, cos, sin, g["ckv"][idx], nseq=nseq)
                    g["g"].append(cg)
            torch.cuda.current_stream().wait_stream(side)
            self._graphs = {key: g}
        g["x"].copy_(X)
        g["e0"].copy_(e0)
        for idx in range(len(self.blocks)):
            g["ckv"][idx].copy_(ckv_fn(idx))
        return g

    # ------------------------------------------------------------------ forward (reference contract)
    @torch.no_grad()
    def forward(self, x, t, context, y=None, freqs=None, pipeline=None, current_step_no=0, real_step_no=0, x_id=0,
                max_steps=0, callback=None, **kwargs):
        """x: list of [B,Cin,T,H,W] (CONSUMED, model.py:1558-1559); t: [1]; context: list of [1|B, Lt, text_dim];
        y: [Cy,T,H,W] (i2v2_2); freqs: (cos, sin) fp32 [L,128].  Returns one fp32 [B,16,T,H,W] per list entry,
        or [None]*n if `pipeline._interrupt` is raised (polled once per block, model.py:1995-1998)."""
        if not self._ready:
            raise RuntimeError("WanModel: load_state_dict() / init_synthetic() must be called before forward()")
        for k in _UNSUPPORTED_KWARGS:
            if kwargs.get(k) is not None:
                raise NotImplementedError(f"WanModel.forward: `{k}` belongs to a conditioning variant outside the t2v/i2v2_2 hot path")
        if torch.is_tensor(t) and t.numel() > 1:
            raise NotImplementedError("per-frame timesteps (diffusion forcing) are outside the t2v/i2v2_2 hot path")
        x_list = list(x)
        if isinstance(x, list):
            x.clear()
        n = len(x_list)
        ctx_list = list(context)
        if len(ctx_list) != n:
            ctx_list = (ctx_list * n)[:n]
        B, _, T, H, W = x_list[0].shape
        thw = (T, H, W)
        shared_state["embed_sizes"] = (T, H // 2, W // 2)
        shared_state["step_no"] = current_step_no
        shared_state["max_steps"] = max_steps
        cos, sin = self._freqs(freqs, thw)
        e, e0 = self._time(t)
        yd = None if y is None else y.to(self.device, f32).contiguous()
        # text embedding per entry (model.py:1856); optionally cached across steps
        ctx_emb, ctx_keys = [], []
        for c in ctx_list:
            key = _PromptCache.key(c) if self.cache_context else None
            ctx_keys.append(key)
            emb = self._prompts.get_emb(key) if key is not None else None
            if emb is None:
                emb = self._text(c[0] if c.dim() == 3 else c)
                if key is not None:
                    self._prompts.put_emb(key, emb, c)
            ctx_emb.append(emb)
        # patch embedding: the fp32 residual streams of ALL (entry, batch item) pairs stacked along the rows of one tensor [m L, D]: the
        # row-wise kernels (LN, GEMMs, RMSNorm+RoPE, FFN) then run once per block for the whole CFG pair -- half the launches and twice
        # the tiles per launch (wave quantisation of the short sequences: 168 -> 336 tiles on 148 SMs at the 1.3B / 480p sizes) -- and
        # the two attentions take the sequence count.  Row-wise arithmetic is unchanged, so results are bit-identical to one stream at a time.
        owners = [(i, b) for i, xi in enumerate(x_list) for b in range(xi.shape[0])]
        m = len(owners)
        L = T * (H // 2) * (W // 2)
        X = torch.empty(m * L, self.dim, device=self.device, dtype=f32)
        for j, (i, b) in enumerate(owners):
            xi = x_list[i].to(self.device, f32)
            ops.patch_embed(xi[b].contiguous(), yd, self._g["pe_w"], self._g["pe_b"], self.dim, out=X[j * L:(j + 1) * L])
        del x_list
        if m > 1:
            fk = ("rep", cos.data_ptr(), m)
            rep = self._freqs_cache.get(fk)
            if rep is None:
                rep = (cos.repeat(m, 1), sin.repeat(m, 1))
                self._freqs_cache[fk] = rep
            cos_m, sin_m = rep
        else:
            cos_m, sin_m = cos, sin

        def entry_ckv(i, idx):
            ckv = self._prompts.get_ckv(ctx_keys[i], idx) if ctx_keys[i] is not None else None
            if ckv is None:
                ckv = self._cross_kv(self.blocks[idx], ctx_emb[i])
                if ctx_keys[i] is not None:
                    self._prompts.put_ckv(ctx_keys[i], idx, ckv)             # 2 * L_text * D bf16 per block (10 MB at 14B)
            return ckv

        def stacked_ckv(idx):
            """text K|V of every stream stacked along the rows [m Lt, 2D] (cached with the prompts: step-invariant)."""
            if m == 1:
                return entry_ckv(owners[0][0], idx)
            skey = (tuple(ctx_keys[i] for i, _ in owners), idx) if all(ctx_keys[i] is not None for i, _ in owners) else None
            hit = self._stacked_ckv.get(skey) if skey is not None else None
            if hit is None:
                hit = torch.cat([entry_ckv(i, idx) for i, _ in owners], 0)
                if skey is not None:
                    if len(self._stacked_ckv) >= 2 * len(self.blocks):       # bounded like the prompt cache: at most two prompt sets
                        self._stacked_ckv.clear()
                    self._stacked_ckv[skey] = hit
            return hit

        graphs = self._block_graphs(X, m, e0, stacked_ckv, cos_m, sin_m) if self.use_cuda_graphs else None
        for idx, blk in enumerate(self.blocks):
            shared_state["layer"] = idx
            if callback is not None:
                callback(-1, None, False, True)
            if pipeline is not None and getattr(pipeline, "_interrupt", False):
                return [None] * n
            if graphs is not None:
                graphs["g"][idx].replay()
                continue
            self._block(blk, X, e0, None, cos_m, sin_m, stacked_ckv(idx), nseq=m)
        if graphs is not None:
            X = graphs["x"]
        heads = [self._head(X[j * L:(j + 1) * L], e, thw) for j in range(m)]
        outs, j = [], 0
        for i in range(n):
            nb = sum(1 for o in owners if o[0] == i)
            outs.append(torch.stack(heads[j:j + nb], 0))
            j += nb
        return outs
