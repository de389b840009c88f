"""The glyph-byT5 text encoder of Hunyuan Video 1.5 on the B200 kernels -- a step in front of the denoise path (SURVEY.md section 8f
row 4, Hunyuan side).

The reference builds it as `T5ForConditionalGeneration.from_pretrained("google/byt5-small").get_encoder()` with the glyph colour / font
tokens added to the vocabulary, loads the Glyph-SDXL-v2 checkpoint over it (models/hyvideo/text_encoder/byT5/__init__.py:45-100, 154-205)
and calls it as `byt5_model(text_ids, attention_mask=text_mask.float())[0]` -> [1, 256, 1472] (pipeline_hunyuan_video.py:1033-1039).  That
model is transformers' T5Stack (third-party; classic T5 v1.1: gated GELU-tanh FFN, RMS layer norm, no score scaling, ONE relative position
embedding owned by block 0 and reused by every block) -- the same arithmetic as the reference's own `T5Encoder(shared_pos=True)`
(models/wan/modules/t5.py:268-292); tests/golden/byt5_tiny.npz holds both and they agree bit for bit.  Here it is
`wan2gp_b200.wan.t5.T5Encoder(shared_pos=True)` (embedding rows, T5 RMS norm, 64-wide position-biased attention and the gated product in
csrc/t5_ops.cuh; the linear layers through the tcgen05 GEMMs) behind the Hugging Face call surface, so it drops into
`HunyuanVideoSampler.byt5_model`."""
import torch

from ..wan.t5 import T5Encoder, hf_to_wan_names

BYT5_SMALL = dict(dim=1472, dim_attn=384, dim_ffn=3584, num_heads=6, num_layers=12, num_buckets=32)    # google/byt5-small config.json


class ByT5Encoder(torch.nn.Module):
    """`model(input_ids, attention_mask=mask)[0]` -> fp32 [B, L, dim]; state dict in the transformers T5Stack / T5EncoderModel naming
    (`[encoder.]block.N.layer.0.SelfAttention.q.weight`, ...) or in the reference T5Encoder naming."""

    def __init__(self, vocab_size, dim=1472, dim_attn=384, dim_ffn=3584, num_heads=6, num_layers=12, num_buckets=32, device="cuda"):
        super().__init__()
        self.encoder = T5Encoder(vocab_size, dim, dim_attn, dim_ffn, num_heads, num_layers, num_buckets, shared_pos=True, device=device)
        self.device = torch.device(device)
        self.dtype = torch.bfloat16

    def load_state_dict(self, sd, strict=True, assign=False):
        return self.encoder.load_state_dict(hf_to_wan_names(sd))

    @classmethod
    def from_state_dict(cls, sd, device="cuda", **cfg):
        """Width / depth / vocabulary read off the checkpoint (the glyph checkpoint's vocabulary is byT5's 384 ids + the colour / font tokens)."""
        w = hf_to_wan_names(sd)
        layers = 1 + max(int(k.split(".")[1]) for k in w if k.startswith("blocks."))
        vocab, dim = w["token_embedding.weight"].shape
        dim_attn, dim_ffn = w["blocks.0.attn.q.weight"].shape[0], w["blocks.0.ffn.fc1.weight"].shape[0]
        buckets, heads = w["pos_embedding.embedding.weight"].shape
        m = cls(vocab, dim, dim_attn, dim_ffn, heads, layers, buckets, device=device)
        m.encoder.load_state_dict(w)
        return m

    @torch.no_grad()
    def forward(self, input_ids, attention_mask=None, **unused):
        return (self.encoder(input_ids, attention_mask),)
