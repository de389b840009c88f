"""umT5 text encoder on the GPU (wan2gp_b200/wan/t5.py through the C ABI) against the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import t5_oracle
from wan2gp_b200 import synth

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def rel_l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm())


def _encoder(cfg, sd):
    from wan2gp_b200.wan.t5 import T5Encoder
    enc = T5Encoder(cfg["vocab_size"], cfg["dim"], cfg["dim_attn"], cfg["dim_ffn"], cfg["num_heads"], cfg["num_layers"], cfg["num_buckets"],
                    shared_pos=False, device="cuda")
    enc.load_state_dict(sd)
    return enc


def test_t5_small_matches_oracle_and_reference_fixture():
    g = np.load(os.path.join(GOLDEN, "t5_small.npz"))
    cfg = synth.T5_CONFIGS["t5_small"]
    sd = synth.make_t5_state_dict(cfg, 0)
    L, nv = int(g["length"]), int(g["n_valid"])
    ids, mask = synth.make_t5_inputs(cfg, L, nv, 0)
    out = _encoder(cfg, sd)(ids[None], mask[None])[0]
    emu = t5_oracle.t5_encode(sd, cfg, ids, mask, emulate_bf16=True)
    # same rounding points, different transcendental approximations (tanh.approx GELU in the GEMM epilogue, __expf): two bf16 pipelines
    # agree to about the size of one bf16 rounding per layer (the emulation itself is 7.8e-3 away from the fp32 reference)
    assert rel_l2(out, emu) < 6e-3
    ref = torch.from_numpy(g["out"])                                     # the unmodified reference in fp32
    assert rel_l2(out[:nv], ref[:nv]) < 2e-2                             # rows the pipeline keeps (t5.py:690)


@pytest.mark.parametrize("L,nv", [(512, 377), (512, 512), (96, 5)])
def test_t5_xxl_layer_shape(L, nv):
    """One block at the umT5-XXL widths (4096 / 64 heads / 10240) on the reference's text_len: pair GEMMs at M = 512, 16 query blocks x 64
    heads of the position-biased attention, padding mask."""
    cfg = synth.T5_CONFIGS["t5_1layer_xxl"]
    sd = synth.make_t5_state_dict(cfg, 1)
    ids, mask = synth.make_t5_inputs(cfg, L, nv, 1)
    out = _encoder(cfg, sd)(ids[None], mask[None])[0]
    emu = t5_oracle.t5_encode(sd, cfg, ids, mask, emulate_bf16=True)
    assert torch.isfinite(out).all()
    assert rel_l2(out[:nv], emu[:nv]) < 6e-3


def test_t5_encoder_model_wrapper_is_the_pipeline_callable():
    """T5EncoderModel(texts, device) -> list of [n_tokens, dim] tensors: what WanAny2V._encode_prompt consumes (any2video.py:588-595)."""
    from wan2gp_b200.wan.t5 import T5EncoderModel
    cfg = synth.T5_CONFIGS["t5_small"]
    sd = synth.make_t5_state_dict(cfg, 0)

    def tok(texts):                                                      # stand-in tokenizer: bytes -> ids, padded to text_len
        ids = torch.zeros(len(texts), 64, dtype=torch.long)
        mask = torch.zeros(len(texts), 64, dtype=torch.long)
        for i, t in enumerate(texts):
            b = [1 + (c % 900) for c in t.encode()][:64]
            ids[i, :len(b)] = torch.tensor(b)
            mask[i, :len(b)] = 1
        return ids, mask
    te = T5EncoderModel(text_len=64, device="cuda", state_dict=sd, tokenizer=tok, encoder=_encoder(cfg, sd))
    outs = te(["a red fox", "two words"], "cuda")
    assert [tuple(o.shape) for o in outs] == [(9, 256), (9, 256)]
    ids, mask = tok(["a red fox"])
    emu = t5_oracle.t5_encode(sd, cfg, ids[0], mask[0], emulate_bf16=True)
    assert rel_l2(outs[0], emu[:9]) < 6e-3


# ---------------------------------------------------------------------------------------------- byT5 glyph encoder (shared position bias)
def _byt5(cfg, sd):
    from wan2gp_b200.hyvideo.byt5 import ByT5Encoder
    return ByT5Encoder.from_state_dict(synth.t5_to_hf_t5stack_names(sd, cfg["num_layers"]), device="cuda")


def test_byt5_tiny_matches_oracle_reference_and_transformers_fixture():
    """ByT5Encoder (T5Encoder(shared_pos=True) behind the Hugging Face call surface) against the bf16-emulating oracle and the fixture that
    holds the reference T5Encoder(shared_pos=True) and transformers' T5Stack outputs."""
    g = np.load(os.path.join(GOLDEN, "byt5_tiny.npz"))
    cfg = synth.T5_CONFIGS["byt5_tiny"]
    sd = synth.make_t5_state_dict(cfg, 0)
    L, nv = int(g["length"]), int(g["n_valid"])
    ids, mask = synth.make_t5_inputs(cfg, L, nv, 0)
    out = _byt5(cfg, sd)(ids[None].cuda(), attention_mask=mask[None].float().cuda())[0][0]
    emu = t5_oracle.t5_encode(sd, cfg, ids, mask, emulate_bf16=True)
    r_emu = rel_l2(out[:nv], emu[:nv])
    r_ref, r_hf = rel_l2(out[:nv], torch.from_numpy(g["out"])[:nv]), rel_l2(out[:nv], torch.from_numpy(g["out_hf"])[:nv])
    print(f"byt5_tiny: vs bf16-emulating oracle {r_emu:.3e}; vs reference T5Encoder(shared_pos) {r_ref:.3e}; vs transformers T5Stack {r_hf:.3e}")
    assert r_emu < 6e-3 and r_ref < 2e-2 and r_hf < 2e-2


@pytest.mark.parametrize("L,nv", [(256, 57), (256, 256)])
def test_byt5_small_widths(L, nv):
    """Two blocks at the google/byt5-small widths (1472 / 6 heads x 64 / 3584; the GEMM N and K are not multiples of 256: tail tiles) on the
    reference's byt5_max_length of 256 tokens."""
    cfg = synth.T5_CONFIGS["byt5_2layer"]
    sd = synth.make_t5_state_dict(cfg, 2)
    ids, mask = synth.make_t5_inputs(cfg, L, nv, 2)
    out = _byt5(cfg, sd)(ids[None].cuda(), attention_mask=mask[None].float().cuda())[0][0]
    emu = t5_oracle.t5_encode(sd, cfg, ids, mask, emulate_bf16=True)
    assert tuple(out.shape) == (L, 1472) and torch.isfinite(out).all()
    assert rel_l2(out[:nv], emu[:nv]) < 6e-3
