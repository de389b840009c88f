"""Decoder-only LLM text towers on the GPU (wan2gp_b200/hyvideo/llm.py through the C ABI) against the oracle and the transformers fixtures."""
import os

import numpy as np
import pytest
import torch

from oracle import llm_oracle
from wan2gp_b200 import _lib, synth

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
bf16 = torch.bfloat16


def rel_l2(a, b):
    a, b = a.detach().double().cpu().flatten(), b.detach().double().cpu().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _s():
    return torch.cuda.current_stream().cuda_stream


@pytest.mark.parametrize("L,H,Hk", [(1, 2, 1), (31, 2, 2), (32, 4, 1), (33, 4, 2), (100, 28, 4), (620, 28, 4), (257, 32, 8)])
def test_rope_half_and_causal_gqa_attention_kernels(L, H, Hk):
    """b200_rope_half (in place on the q and k heads of a fused q|k|v buffer) and b200_causal_gqa_attention (head dim 128, grouped kv heads,
    tiles of 32 keys with a partial last tile) against fp64 torch on the same bf16 inputs."""
    g = torch.Generator(device="cuda").manual_seed(L * 131 + H)
    W = (H + 2 * Hk) * 128
    qkv = torch.randn(L, W, device="cuda", generator=g).to(bf16)
    cos, sin = llm_oracle.rope_tables(L, 1e6)
    ref = qkv.double().cpu().clone()
    qk = ref[:, :(H + Hk) * 128].reshape(L, H + Hk, 128)
    ref[:, :(H + Hk) * 128] = llm_oracle.apply_rope(qk, cos.double(), sin.double()).reshape(L, -1)
    cos_d, sin_d = cos.cuda(), sin.cuda()                 # named: a temporary would be freed before the kernel reads it
    _lib.call("b200_rope_half", qkv.data_ptr(), qkv.stride(0), cos_d.data_ptr(), sin_d.data_ptr(), L, H + Hk, _s())
    torch.cuda.synchronize()
    assert rel_l2(qkv[:, :(H + Hk) * 128], ref[:, :(H + Hk) * 128]) < 4e-3 and torch.equal(qkv[:, (H + Hk) * 128:].double().cpu(), ref[:, (H + Hk) * 128:])
    # attention on the roped buffer
    q = qkv[:, :H * 128].double().cpu().reshape(L, H, 128)
    k = qkv[:, H * 128:(H + Hk) * 128].double().cpu().reshape(L, Hk, 128).repeat_interleave(H // Hk, 1)
    v = qkv[:, (H + Hk) * 128:].double().cpu().reshape(L, Hk, 128).repeat_interleave(H // Hk, 1)
    s = torch.einsum("ihc,jhc->hij", q, k) * 128 ** -0.5
    s = s.masked_fill(torch.ones(L, L, dtype=torch.bool).triu(1), float("-inf"))
    want = torch.einsum("hij,jhc->ihc", torch.softmax(s, -1), v).reshape(L, H * 128)
    out = torch.full((L, H * 128), float("nan"), device="cuda", dtype=bf16)
    _lib.call("b200_causal_gqa_attention", qkv.data_ptr(), qkv[:, H * 128:].data_ptr(), qkv[:, (H + Hk) * 128:].data_ptr(), qkv.stride(0), qkv.stride(0),
              out.data_ptr(), out.stride(0), L, H, Hk, 128 ** -0.5, _s())
    torch.cuda.synchronize()
    assert torch.isfinite(out).all() and rel_l2(out, want) < 4e-3          # one bf16 rounding of the output
    assert rel_l2(out[0], v[0].reshape(-1)) < 4e-3                         # row 0 attends to itself only


def _model(cfg, sd, prefix="model.language_model."):
    from wan2gp_b200.hyvideo.llm import LlamaLikeTextModel
    return LlamaLikeTextModel.from_state_dict({prefix + k: v for k, v in sd.items()}, cfg["num_heads"], cfg["num_kv_heads"], cfg["rms_eps"],
                                              cfg["rope_theta"], device="cuda")


@pytest.mark.parametrize("name", ["qwen_tiny", "llama_tiny"])
def test_text_model_matches_oracle_and_transformers_fixture(name):
    g = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    cfg = synth.LLM_CONFIGS[name]
    sd = synth.make_llm_state_dict(cfg, 0)
    L, nv = int(g["length"]), int(g["n_valid"])
    ids, mask = synth.make_llm_inputs(cfg, L, nv, 0)
    out = _model(cfg, sd)(input_ids=ids[None].cuda(), attention_mask=mask[None].cuda(), output_hidden_states=True)
    emu = llm_oracle.llm_hidden_states(sd, cfg, ids, nv, emulate_bf16=True)
    ref = torch.from_numpy(g["hidden_states"])
    assert len(out.hidden_states) == cfg["num_layers"] + 1
    for i in (1, -3, -1):
        r_emu, r_ref = rel_l2(out.hidden_states[i][0, :nv], emu[i][:nv]), rel_l2(out.hidden_states[i][0, :nv], ref[i])
        print(f"{name} hidden_states[{i}]: vs bf16-emulating oracle {r_emu:.3e}; vs transformers fp32 {r_ref:.3e}")
        assert r_emu < 6e-3 and r_ref < 2e-2
    assert float(out.hidden_states[-3][0, nv:].abs().max()) == 0.0 and torch.equal(out.last_hidden_state, out.hidden_states[-1])


def test_qwen25_vl_7b_layer_widths():
    """Two layers at the Qwen2.5-VL-7B widths (3584 / 28 q heads : 4 kv heads / 18944) on a prompt of the reference's length (text_len 512 +
    the template's ~108 tokens, right padded): pair GEMMs with bias / SiLU epilogues, 78 x 28 attention blocks."""
    cfg = synth.LLM_CONFIGS["qwen_2layer_7b"]
    sd = synth.make_llm_state_dict(cfg, 1)
    L, nv = 620, 457
    ids, mask = synth.make_llm_inputs(cfg, L, nv, 1)
    out = _model(cfg, sd, prefix="model.")(input_ids=ids[None].cuda(), attention_mask=mask[None].cuda(), output_hidden_states=True)
    emu = llm_oracle.llm_hidden_states(sd, cfg, ids, nv, emulate_bf16=True)
    for i in (1, 2):
        assert torch.isfinite(out.hidden_states[i]).all()
        assert rel_l2(out.hidden_states[i][0, :nv], emu[i][:nv]) < 6e-3


def test_reference_text_encoder_protocol_with_swapped_model(monkeypatch):
    """What the Hunyuan handler does: the reference's TextEncoder object keeps its tokenizer / template / crop logic and only `.model` is
    replaced.  A minimal wrapper with the reference's encode() arithmetic (text_encoder_1_5.py:470-499: hidden_states[-(skip + 1)], crop_start)
    drives the swapped model; the result feeds HunyuanVideoSampler through the text2tokens / encode protocol."""
    import types

    from tests.test_hy_plugin_cpu import hy_kwargs, make_pipeline
    pipe_obj, pipe, hcfg, _, _ = make_pipeline("b200_hunyuan_1_5_t2v", device="cuda", monkeypatch=monkeypatch, vae_tiling=False)
    cfg = dict(synth.LLM_CONFIGS["qwen_tiny"], hidden_size=hcfg["text_states_dim"] if hcfg["text_states_dim"] % 128 == 0 else 256)
    sd = synth.make_llm_state_dict(cfg, 3)
    model = _model(cfg, sd)
    proj = torch.randn(cfg["hidden_size"], hcfg["text_states_dim"], generator=torch.Generator().manual_seed(0)) * cfg["hidden_size"] ** -0.5

    class RefLikeTextEncoder:                       # the reference wrapper's control flow around `self.model`
        max_length, crop_start, skip = 24, 5, 2

        def __init__(self):
            self.model = model

        def text2tokens(self, prompts, data_type="video", max_length=None):
            L = self.max_length + self.crop_start
            ids, mask = torch.zeros(len(prompts), L, dtype=torch.long), torch.zeros(len(prompts), L, dtype=torch.long)
            for i, p in enumerate(prompts):
                b = [7, 8, 9, 10, 11] + [12 + c % 250 for c in p.encode()][:self.max_length]
                ids[i, :len(b)], mask[i, :len(b)] = torch.tensor(b), 1
            return {"input_ids": ids, "attention_mask": mask}

        def encode(self, tok, data_type="video", device=None, is_uncond=False):
            out = self.model(input_ids=tok["input_ids"].to(self.model.device), attention_mask=tok["attention_mask"].to(self.model.device), output_hidden_states=True)
            h = out.hidden_states[-(self.skip + 1)][:, self.crop_start:]
            return types.SimpleNamespace(hidden_state=h @ proj.to(h.device), attention_mask=tok["attention_mask"][:, self.crop_start:])
    te = RefLikeTextEncoder()
    pipe_obj.text_encoder = te
    tok = te.text2tokens(["a red fox"])
    got = te.encode(tok)
    nv = int(tok["attention_mask"].sum())
    emu = llm_oracle.llm_hidden_states(sd, cfg, tok["input_ids"][0], nv, emulate_bf16=True)[-3]
    assert rel_l2(got.hidden_state[0, :nv - 5], emu[5:nv] @ proj) < 6e-3 and int(got.attention_mask.sum()) == nv - 5
    out = pipe_obj.generate(**hy_kwargs(sampling_steps=2, seed=4))
    assert tuple(out.shape) == (3, 5, 32, 48) and torch.isfinite(out).all()
