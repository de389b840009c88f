"""umT5 text encoder on the B200 kernels -- the step in front of the denoise path.

Mirrors /root/reference/models/wan/modules/t5.py: `T5Encoder` (:268-292; the umT5 layout `shared_pos=False`: every block owns its
relative position embedding, :165-188, and the classic T5 / byT5 layout `shared_pos=True`: one embedding for all blocks, :275-276, 286-287)
with the reference's state-dict names, and `T5EncoderModel` (:632-690), the object
`WanAny2V` holds as `self.text_encoder` and calls as `text_encoder([prompt], device) -> [ctx [n_tokens, 4096]]` (any2video.py:125, :588-589).

One prompt = one sequence of text_len (512) token ids.  Per block: T5LayerNorm -> fused q|k|v GEMM -> 64-wide attention with the block's
position bias and the padding mask (no score scaling) -> o GEMM accumulated into the fp32 residual stream; T5LayerNorm -> gate GEMM with
the GELU-tanh epilogue and fc1 GEMM -> product -> fc2 GEMM accumulated into the residual stream.  The reference runs the block in bf16
end to end; here the residual stream and the norms are fp32 and only GEMM operands are bf16 (oracle/t5_oracle.py, emulate_bf16)."""
import math

import torch

from .. import _lib, ops

bf16, f32 = torch.bfloat16, torch.float32
_ACT_GELU_TANH = 1


def _s():
    return torch.cuda.current_stream().cuda_stream


def relative_position_bucket(rel_pos, num_buckets, max_dist=128):
    """T5RelativeEmbedding._relative_position_bucket (t5.py:246-265), bidirectional; rel_pos = key index - query index (int64)."""
    nb = num_buckets // 2
    buckets = (rel_pos > 0).long() * nb
    rel_pos = rel_pos.abs()
    max_exact = nb // 2
    large = max_exact + (torch.log(rel_pos.float() / max_exact) / math.log(max_dist / max_exact) * (nb - max_exact)).long()
    large = torch.min(large, torch.full_like(large, nb - 1))
    return buckets + torch.where(rel_pos < max_exact, rel_pos, large)


def relative_bias_table(emb_weight, L, num_buckets):
    """[num_buckets, heads] embedding -> fp32 [heads, 2L-1]: the bias of key offset j - i = o at column o + L - 1 (what T5RelativeEmbedding
    .forward, t5.py:232-244, expands to [heads, L, L])."""
    rel = torch.arange(-(L - 1), L, device=emb_weight.device)
    return emb_weight.float()[relative_position_bucket(rel, num_buckets)].t().contiguous()


class T5Encoder(torch.nn.Module):
    """Same constructor arguments and state-dict names as the reference class.  `shared_pos=False` = umT5 (Wan, t5.py:469);
    `shared_pos=True` = classic T5 v1.1 / byT5 (the glyph encoder of Hunyuan Video 1.5): `pos_embedding.embedding.weight` serves every block."""

    def __init__(self, vocab, dim, dim_attn, dim_ffn, num_heads, num_layers, num_buckets, shared_pos=False, dropout=0.1, device="cuda"):
        super().__init__()
        self.shared_pos = bool(shared_pos)
        if dim_attn // num_heads != 64:
            raise NotImplementedError("T5Encoder: head dim 64 only (umT5-XXL: 4096 / 64)")
        self.vocab_size, self.dim, self.dim_attn, self.dim_ffn = int(vocab), dim, dim_attn, dim_ffn
        self.num_heads, self.num_layers, self.num_buckets = num_heads, num_layers, num_buckets
        self.device = torch.device(device)
        self._ready = False
        self._bias_cache = {}

    def load_state_dict(self, sd, strict=True, assign=False):
        dev = self.device
        g = lambda k: sd[k].detach()
        self.table = g("token_embedding.weight").to(dev, bf16).contiguous()
        self.norm_w = g("norm.weight").to(dev, f32).contiguous()
        shared = g("pos_embedding.embedding.weight").to(dev, f32).contiguous() if self.shared_pos else None
        self.blocks = []
        for i in range(self.num_layers):
            b = f"blocks.{i}."
            self.blocks.append(dict(
                n1=g(b + "norm1.weight").to(dev, f32).contiguous(), n2=g(b + "norm2.weight").to(dev, f32).contiguous(),
                wqkv=torch.cat([g(b + "attn.q.weight"), g(b + "attn.k.weight"), g(b + "attn.v.weight")]).to(dev, bf16).contiguous(),
                wo=g(b + "attn.o.weight").to(dev, bf16).contiguous(),
                wg=g(b + "ffn.gate.0.weight").to(dev, bf16).contiguous(), w1=g(b + "ffn.fc1.weight").to(dev, bf16).contiguous(),
                w2=g(b + "ffn.fc2.weight").to(dev, bf16).contiguous(),
                pos=shared if self.shared_pos else g(b + "pos_embedding.embedding.weight").to(dev, f32).contiguous()))
        self._bias_cache = {}
        self._ready = True
        return torch.nn.modules.module._IncompatibleKeys([], [])

    def _bias(self, i, L):
        key = (0 if self.shared_pos else i, L)
        if key not in self._bias_cache:
            self._bias_cache[key] = relative_bias_table(self.blocks[i]["pos"], L, self.num_buckets)
        return self._bias_cache[key]

    def _rms(self, x, w, out_fp32=False):
        L, D = x.shape
        out = torch.empty(L, D, device=x.device, dtype=f32 if out_fp32 else bf16)
        _lib.call("b200_t5_rmsnorm", x.data_ptr(), w.data_ptr(), out.data_ptr(), int(out_fp32), L, D, 1e-6, _s())
        return out

    @torch.no_grad()
    def encode_one(self, ids, n_valid):
        """ids int64 [L] on the device, the first n_valid are tokens (the rest padding) -> fp32 [L, dim]."""
        if not self._ready:
            raise RuntimeError("T5Encoder: load_state_dict() must be called first")
        L, D, DA = ids.numel(), self.dim, self.dim_attn
        x = torch.empty(L, D, device=self.device, dtype=f32)
        _lib.call("b200_embed_rows", ids.data_ptr(), self.table.data_ptr(), 1, x.data_ptr(), L, D, _s())
        for i, b in enumerate(self.blocks):
            qkv = ops.gemm(self._rms(x, b["n1"]), b["wqkv"])                                         # [L, 3 DA] bf16
            att = torch.empty(L, DA, device=self.device, dtype=bf16)
            _lib.call("b200_t5_attention", qkv.data_ptr(), qkv[:, DA:].data_ptr(), qkv[:, 2 * DA:].data_ptr(), qkv.stride(0),
                      self._bias(i, L).data_ptr(), att.data_ptr(), att.stride(0), L, self.num_heads, int(n_valid), _s())
            ops.gemm(att, b["wo"], out=x, accumulate=True)                                           # x += attn(norm1(x))   (t5.py:184)
            xn = self._rms(x, b["n2"])
            hg = ops.gemm(xn, b["wg"], act=_ACT_GELU_TANH)
            hu = ops.gemm(xn, b["w1"])
            _lib.call("b200_mul_bf16", hu.data_ptr(), hg.data_ptr(), hg.data_ptr(), hg.numel(), _s())   # fc1(x) * gelu(gate(x)), in place
            ops.gemm(hg, b["w2"], out=x, accumulate=True)                                            # x += ffn(norm2(x))    (t5.py:185)
        return self._rms(x, self.norm_w, out_fp32=True)

    @torch.no_grad()
    def forward(self, ids, mask=None):
        """ids [B, L] int64, mask [B, L] (1 = token) -> [B, L, dim] fp32 (T5Encoder.forward, t5.py:282-292; eval mode: dropout is identity)."""
        ids = ids.to(self.device)
        outs = []
        for bi in range(ids.shape[0]):
            n_valid = int(mask[bi].gt(0).sum()) if mask is not None else ids.shape[1]
            if mask is not None and not bool((mask[bi, :n_valid] > 0).all()):
                raise NotImplementedError("T5Encoder: the mask must be a prefix of ones (tokenizer padding), as the reference tokenizer produces")
            outs.append(self.encode_one(ids[bi].contiguous(), max(n_valid, 1)))
        return torch.stack(outs)


def umt5_xxl_encoder(device="cuda"):
    """umt5_xxl(encoder_only=True) of t5.py:459-472."""
    return T5Encoder(256384, 4096, 4096, 10240, 64, 24, 32, shared_pos=False, device=device)


class T5EncoderModel:
    """The object any2video.py:125 builds: `T5EncoderModel(text_len, dtype, device, checkpoint_path, tokenizer_path)`, called as
    `model(texts, device) -> [context [n_tokens_i, 4096]]` (t5.py:683-690).  `state_dict` (reference names) or `checkpoint_path` (a
    safetensors / torch file in the Wan or the Hugging Face umT5 naming -- the latter through `hf_to_wan_names`) supply the weights;
    the tokenizer is transformers' AutoTokenizer on `tokenizer_path` (the reference's HuggingfaceTokenizer wraps the same class,
    tokenizers.py), or any callable `tokenizer(texts) -> (ids [B, L], mask [B, L])`."""

    def __init__(self, text_len, dtype=bf16, device="cuda", checkpoint_path=None, tokenizer_path=None, state_dict=None, tokenizer=None,
                 encoder=None):
        self.text_len, self.dtype, self.device = text_len, dtype, torch.device(device)
        self.model = encoder if encoder is not None else umt5_xxl_encoder(self.device)
        if state_dict is None and checkpoint_path is not None:
            if str(checkpoint_path).endswith(".safetensors"):
                from safetensors.torch import load_file
                state_dict = load_file(checkpoint_path)
            else:
                state_dict = torch.load(checkpoint_path, map_location="cpu")
        if state_dict is not None:
            self.model.load_state_dict(hf_to_wan_names(state_dict))
        self.tokenizer = tokenizer
        if tokenizer is None and tokenizer_path is not None:
            from transformers import AutoTokenizer
            tok = AutoTokenizer.from_pretrained(tokenizer_path)

            def _tok(texts):
                texts = [clean_prompt(t) for t in ([texts] if isinstance(texts, str) else texts)]     # HuggingfaceTokenizer(clean='whitespace')
                enc = tok(texts, return_tensors="pt", padding="max_length", truncation=True, max_length=self.text_len, add_special_tokens=True)
                return enc.input_ids, enc.attention_mask
            self.tokenizer = _tok

    def __call__(self, texts, device=None):
        if self.tokenizer is None:
            raise RuntimeError("T5EncoderModel: no tokenizer (pass tokenizer_path or a tokenizer callable)")
        ids, mask = self.tokenizer(texts)
        ctx = self.model(ids, mask)
        seq_lens = mask.gt(0).sum(dim=1).long()
        return [u[:int(v)] for u, v in zip(ctx, seq_lens)]


def clean_prompt(text):
    """HuggingfaceTokenizer._clean with clean='whitespace' (tokenizers.py:12-23, :77-79): ftfy.fix_text when ftfy is installed (WanGP
    requires it), two rounds of html.unescape, runs of whitespace collapsed to one space."""
    import html
    import re
    try:
        import ftfy
        text = ftfy.fix_text(text)
    except ImportError:
        pass
    text = html.unescape(html.unescape(text)).strip()
    return re.sub(r"\s+", " ", text).strip()


def hf_to_wan_names(sd):
    """Hugging Face UMT5EncoderModel names -> the reference's names (what convert_umt5_encoder_to_wan_format, t5.py:497-621, produces);
    a dict that already uses the reference's names is returned unchanged."""
    if "token_embedding.weight" in sd:
        return sd
    import re
    if not any(k.startswith("encoder.") for k in sd):          # a bare T5Stack (`T5ForConditionalGeneration.get_encoder()`, byT5): no prefix
        sd = {"encoder." + k: v for k, v in sd.items()}
    bias_blocks = [k for k in sd if k.endswith("SelfAttention.relative_attention_bias.weight")]
    shared = len(bias_blocks) == 1 and sum(1 for k in sd if k.endswith("layer.0.SelfAttention.q.weight")) > 1
    out = {}
    rules = [(r"^encoder\.final_layer_norm\.weight$", "norm.weight"),
             (r"^encoder\.block\.(\d+)\.layer\.0\.layer_norm\.weight$", "blocks.{}.norm1.weight"),
             (r"^encoder\.block\.(\d+)\.layer\.1\.layer_norm\.weight$", "blocks.{}.norm2.weight"),
             (r"^encoder\.block\.(\d+)\.layer\.0\.SelfAttention\.([qkvo])\.weight$", "blocks.{}.attn.{}.weight"),
             (r"^encoder\.block\.(\d+)\.layer\.0\.SelfAttention\.relative_attention_bias\.weight$", "blocks.{}.pos_embedding.embedding.weight"),
             (r"^encoder\.block\.(\d+)\.layer\.1\.DenseReluDense\.wi_0\.weight$", "blocks.{}.ffn.gate.0.weight"),
             (r"^encoder\.block\.(\d+)\.layer\.1\.DenseReluDense\.wi_1\.weight$", "blocks.{}.ffn.fc1.weight"),
             (r"^encoder\.block\.(\d+)\.layer\.1\.DenseReluDense\.wo\.weight$", "blocks.{}.ffn.fc2.weight")]
    for k, v in sd.items():
        if k in ("shared.weight", "encoder.embed_tokens.weight"):
            out.setdefault("token_embedding.weight", v)
            continue
        if shared and k == bias_blocks[0]:                       # classic T5: block 0's bias is every block's (T5Stack passes position_bias on)
            out["pos_embedding.embedding.weight"] = v
            continue
        for pat, repl in rules:
            m = re.match(pat, k)
            if m:
                out[repl.format(*m.groups())] = v
                break
    return out
