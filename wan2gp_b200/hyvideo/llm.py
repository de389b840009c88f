"""Decoder-only LLM text towers on the B200 kernels -- the text encoders in front of the Hunyuan denoise path (SURVEY.md section 8f row 4,
Hunyuan side): the language model of Qwen2.5-VL-7B-Instruct (Hunyuan Video 1.5) and of llava-llama-3-8b (HunyuanVideo 1.0).

The reference loads them through transformers (`Qwen2_5_VLForConditionalGeneration`, models/hyvideo/text_encoder/text_encoder_1_5.py:86-117;
`LlavaForConditionalGeneration`, text_encoder/__init__.py) and its `TextEncoder.encode` calls
`self.model(input_ids=..., attention_mask=..., output_hidden_states=True)` and keeps `outputs.hidden_states[-(skip + 1)]` with skip = 2
(text_encoder_1_5.py:470-482, hunyuan.py:433).  `LlamaLikeTextModel` is that callable: same keyword arguments, a result with
`.hidden_states` (a tuple of num_layers + 1 tensors [B, L, D]: embeddings, layer outputs, the LAST one after the final norm) and
`.last_hidden_state`, `.device` / `.dtype`, and `final_layer_norm` -- so `text_encoder.model = LlamaLikeTextModel.from_state_dict(...)` swaps
it into the reference's own `TextEncoder` object, whose tokenizer, prompt templates and crop logic keep running unchanged
(plugin/models/b200_hunyuan_handler.py).

Per layer: RMSNorm (b200_t5_rmsnorm) -> fused q|k|v GEMM (+bias) -> rotate-half RoPE in place (b200_rope_half) -> causal grouped-query
attention, head dim 128 (b200_causal_gqa_attention) -> o GEMM accumulated into the fp32 residual stream; RMSNorm -> gate GEMM with the SiLU
epilogue and up GEMM -> product (b200_mul_bf16) -> down GEMM accumulated into the residual stream.  Text prompts only (no image tokens):
the multimodal RoPE of Qwen2.5-VL has three equal position axes then and reduces to the plain rotary embedding.  With right padding
(`padding_side="right"`, text_encoder_1_5.py:311-316) and a causal mask no valid row sees a padded key: a sequence is encoded on its valid
prefix; padded rows (don't-care in transformers, cropped by the attention mask downstream) are returned as zeros.

The reference runs the tower in bf16 end to end; here the residual stream and the norms are fp32 and GEMM operands are bf16
(oracle/llm_oracle.py, emulate_bf16)."""
import types

import torch

from .. import _lib, ops

bf16, f32 = torch.bfloat16, torch.float32
_ACT_SILU = 2

QWEN25_VL_7B = dict(hidden_size=3584, intermediate_size=18944, num_layers=28, num_heads=28, num_kv_heads=4, rms_eps=1e-6, rope_theta=1e6)
LLAMA3_8B = dict(hidden_size=4096, intermediate_size=14336, num_layers=32, num_heads=32, num_kv_heads=8, rms_eps=1e-5, rope_theta=5e5)


def _s():
    return torch.cuda.current_stream().cuda_stream


def rope_tables(n, theta, head_dim=128, device="cpu"):
    """transformers RotaryEmbedding for positions 0..n-1: (cos, sin) fp32 [n, head_dim / 2] (the library concatenates two copies)."""
    inv_freq = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64).float() / head_dim))
    freqs = torch.arange(n, dtype=f32)[:, None] * inv_freq[None, :]
    return freqs.cos().to(device).contiguous(), freqs.sin().to(device).contiguous()


def strip_prefix(sd):
    """Language-model weights of a transformers checkpoint under any of its prefixes (`model.language_model.` for Qwen2.5-VL since
    transformers 4.52, `model.` before / for Llama, `language_model.model.` inside Llava) -> bare names (`layers.0...`, `embed_tokens`, `norm`)."""
    emb = [k for k in sd if k.endswith("embed_tokens.weight") and "visual" not in k and "vision" not in k]
    if not emb:
        raise KeyError("no `embed_tokens.weight` in the state dict: not a transformers language-model checkpoint")
    pre = min(emb, key=len)[:-len("embed_tokens.weight")]
    return {k[len(pre):]: v for k, v in sd.items() if k.startswith(pre) and k[len(pre):].startswith(("layers.", "embed_tokens.", "norm."))}


class LlamaLikeTextModel(torch.nn.Module):
    """Llama-architecture decoder stack (Qwen2: q/k/v biases; Llama-3: none) with head dim 128, used as a text ENCODER: no KV cache, no
    logits, every layer's hidden state is returned."""

    def __init__(self, vocab_size, hidden_size, intermediate_size, num_layers, num_heads, num_kv_heads, rms_eps=1e-6, rope_theta=1e6,
                 device="cuda"):
        super().__init__()
        if hidden_size % 8 or num_heads % num_kv_heads:
            raise ValueError("LlamaLikeTextModel: hidden size / head grouping")
        self.vocab_size, self.hidden_size, self.intermediate_size = int(vocab_size), hidden_size, intermediate_size
        self.num_layers, self.num_heads, self.num_kv_heads = num_layers, num_heads, num_kv_heads
        self.rms_eps, self.rope_theta = float(rms_eps), float(rope_theta)
        self._device = torch.device(device)
        self._ready = False
        self._rope = {}

    @property
    def device(self):
        return self._device

    @property
    def dtype(self):
        return bf16

    def load_state_dict(self, sd, strict=True, assign=False):
        sd, dev = strip_prefix(sd), self._device
        g = lambda k: sd[k].detach()
        self.table = g("embed_tokens.weight").to(dev, bf16).contiguous()
        self.norm_w = g("norm.weight").to(dev, f32).contiguous()
        self.layers = []
        for i in range(self.num_layers):
            p = f"layers.{i}."
            names = [p + f"self_attn.{n}_proj" for n in "qkv"]
            bias = None
            if names[0] + ".bias" in sd:
                bias = torch.cat([g(n + ".bias") for n in names]).to(dev, f32).contiguous()
            self.layers.append(dict(
                n1=g(p + "input_layernorm.weight").to(dev, f32).contiguous(), n2=g(p + "post_attention_layernorm.weight").to(dev, f32).contiguous(),
                wqkv=torch.cat([g(n + ".weight") for n in names]).to(dev, bf16).contiguous(), bqkv=bias,
                wo=g(p + "self_attn.o_proj.weight").to(dev, bf16).contiguous(), wg=g(p + "mlp.gate_proj.weight").to(dev, bf16).contiguous(),
                wu=g(p + "mlp.up_proj.weight").to(dev, bf16).contiguous(), wd=g(p + "mlp.down_proj.weight").to(dev, bf16).contiguous()))
        self._ready = True
        return torch.nn.modules.module._IncompatibleKeys([], [])

    @classmethod
    def from_state_dict(cls, sd, num_heads, num_kv_heads, rms_eps=1e-6, rope_theta=1e6, device="cuda"):
        """Widths, depth and vocabulary read off the checkpoint; the head grouping, epsilon and RoPE base come from its config.json."""
        w = strip_prefix(sd)
        layers = 1 + max(int(k.split(".")[1]) for k in w if k.startswith("layers."))
        vocab, dim = w["embed_tokens.weight"].shape
        m = cls(vocab, dim, w["layers.0.mlp.gate_proj.weight"].shape[0], layers, num_heads, num_kv_heads, rms_eps, rope_theta, device=device)
        if w["layers.0.self_attn.q_proj.weight"].shape[0] != num_heads * 128 or w["layers.0.self_attn.k_proj.weight"].shape[0] != num_kv_heads * 128:
            raise ValueError("LlamaLikeTextModel: head dim 128 only (q / k projection rows do not match heads x 128)")
        m.load_state_dict(w)
        return m

    def _rms(self, x, w, out_fp32=False):
        L, D = x.shape
        out = torch.empty(L, D, device=x.device, dtype=f32 if out_fp32 else bf16)
        _lib.call("b200_t5_rmsnorm", x.data_ptr(), w.data_ptr(), out.data_ptr(), int(out_fp32), L, D, self.rms_eps, _s())
        return out

    def final_layer_norm(self, x):
        """`text_encoder.model.final_layer_norm` of the reference wrapper (text_encoder_1_5.py:102-103); x [..., D]."""
        shape = x.shape
        return self._rms(x.reshape(-1, shape[-1]).to(self._device, f32).contiguous(), self.norm_w, out_fp32=True).reshape(shape)

    @torch.no_grad()
    def hidden_states_one(self, ids):
        """ids int64 [n] on the device (valid tokens only) -> list of num_layers + 1 fp32 [n, D] tensors."""
        if not self._ready:
            raise RuntimeError("LlamaLikeTextModel: load_state_dict() must be called first")
        n, D, H, Hk = ids.numel(), self.hidden_size, self.num_heads, self.num_kv_heads
        if n not in self._rope:
            self._rope = {n: rope_tables(n, self.rope_theta, 128, self._device)}
        cos, sin = self._rope[n]
        x = torch.empty(n, D, device=self._device, dtype=f32)
        _lib.call("b200_embed_rows", ids.data_ptr(), self.table.data_ptr(), 1, x.data_ptr(), n, D, _s())
        hs = [x.clone()]
        nq, nkv = H * 128, Hk * 128
        for b in self.layers:
            qkv = ops.gemm(self._rms(x, b["n1"]), b["wqkv"], bias=b["bqkv"])                         # [n, (H + 2 Hk) 128] bf16
            _lib.call("b200_rope_half", qkv.data_ptr(), qkv.stride(0), cos.data_ptr(), sin.data_ptr(), n, H + Hk, _s())
            att = torch.empty(n, nq, device=self._device, dtype=bf16)
            _lib.call("b200_causal_gqa_attention", qkv.data_ptr(), qkv[:, nq:].data_ptr(), qkv[:, nq + nkv:].data_ptr(), qkv.stride(0), qkv.stride(0),
                      att.data_ptr(), att.stride(0), n, H, Hk, 128 ** -0.5, _s())
            ops.gemm(att, b["wo"], out=x, accumulate=True)                                           # x += o_proj(attention)
            xn = self._rms(x, b["n2"])
            hg = ops.gemm(xn, b["wg"], act=_ACT_SILU)
            hu = ops.gemm(xn, b["wu"])
            _lib.call("b200_mul_bf16", hu.data_ptr(), hg.data_ptr(), hg.data_ptr(), hg.numel(), _s())   # silu(gate(x)) * up(x), in place
            ops.gemm(hg, b["wd"], out=x, accumulate=True)                                            # x += down(...)
            hs.append(x.clone())
        hs[-1] = self._rms(x, self.norm_w, out_fp32=True)                                            # transformers norms the last entry
        return hs

    @torch.no_grad()
    def forward(self, input_ids=None, attention_mask=None, output_hidden_states=False, **unused):
        """transformers call surface: -> namespace(last_hidden_state [B, L, D], hidden_states tuple of [B, L, D] or None)."""
        ids = input_ids.to(self._device)
        B, L = ids.shape
        per = []
        for bi in range(B):
            n = int(attention_mask[bi].gt(0).sum()) if attention_mask is not None else L
            if attention_mask is not None and not bool((attention_mask[bi, :n] > 0).all()):
                raise NotImplementedError("LlamaLikeTextModel: the mask must be a prefix of ones (right padding), as the reference tokenizers produce")
            hs = self.hidden_states_one(ids[bi, :max(n, 1)].contiguous())
            per.append([torch.cat([h, h.new_zeros(L - h.shape[0], h.shape[1])]) for h in hs])
        stacked = tuple(torch.stack([p[i] for p in per]) for i in range(self.num_layers + 1))
        return types.SimpleNamespace(last_hidden_state=stacked[-1], hidden_states=stacked if output_hidden_states else None)
