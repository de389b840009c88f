"""Decoder-only LLM text towers in front of the Hunyuan path (Hunyuan side): the oracle against fixtures produced by
transformers' own Qwen2_5_VLTextModel / LlamaModel, the host-side pieces of wan2gp_b200/hyvideo/llm.py, and a lane-by-lane numpy
restatement of the attention kernel's loop structure (csrc/llm_ops.cuh) against a plain softmax."""
import os

import numpy as np
import pytest
import torch

from oracle import llm_oracle
from tests.test_host_dryrun_cpu import stub_abi  # noqa: F401  (fixture)
from wan2gp_b200 import synth

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name", ["qwen_tiny", "llama_tiny"])
def test_oracle_matches_transformers_fixture(name):
    """tests/golden/{qwen,llama}_tiny.npz: every hidden state of transformers' Qwen2_5_VLTextModel (multimodal RoPE sections, q/k/v biases,
    grouped-query attention 2:1) / LlamaModel (no biases, 4:2) on the valid rows of a right-padded sequence (oracle/gen_golden.py)."""
    g = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    cfg = synth.LLM_CONFIGS[name]
    sd = synth.make_llm_state_dict(cfg, int(g["seed"]))
    nv = int(g["n_valid"])
    ids, mask = synth.make_llm_inputs(cfg, int(g["length"]), nv, int(g["seed"]))
    hs = llm_oracle.llm_hidden_states(sd, cfg, ids, nv)
    ref = torch.from_numpy(g["hidden_states"])
    assert len(hs) == cfg["num_layers"] + 1 == ref.shape[0]
    for i, (h, r) in enumerate(zip(hs, ref)):
        assert float((h[:nv] - r).norm() / r.norm()) < 1e-6, i
        assert float(h[nv:].abs().max()) == 0.0
    # the prefix property the product relies on: the valid rows do not depend on what follows them
    hs_short = llm_oracle.llm_hidden_states(sd, cfg, ids[:nv], nv)
    assert torch.equal(hs_short[-3], hs[-3][:nv])


def test_attention_kernel_loop_structure_in_numpy():
    """causal_gqa_attention_kernel restated lane by lane (tiles of 32 keys, lane l owns key j0 + l for the score and output dims 4l..4l+3,
    online softmax in base 2 with a pre-scaled query, probabilities broadcast by 'shuffle') equals softmax(q k^T / sqrt(d)) v."""
    rng = np.random.default_rng(0)
    L, H, Hk, d = 77, 4, 2, 128
    q, k, v = rng.standard_normal((L, H, d)), rng.standard_normal((L, Hk, d)), rng.standard_normal((L, Hk, d))
    scale_log2e = d ** -0.5 * 1.4426950408889634
    out = np.zeros((L, H, d))
    lanes = np.arange(32)
    for h in range(H):
        hk = h // (H // Hk)
        for i in range(L):
            qs = q[i, h] * scale_log2e
            m, l, acc = -np.inf, 0.0, np.zeros((32, 4))                      # acc[lane] = output dims 4 lane .. 4 lane + 3
            for j0 in range(0, i + 1, 32):
                j = j0 + lanes
                valid = j <= i
                s = np.where(valid, (k[np.minimum(j, L - 1), hk] * qs).sum(-1), -np.inf)
                m_new = max(m, s.max())
                corr = np.exp2(m - m_new)
                p = np.where(valid, np.exp2(s - m_new), 0.0)
                l = l * corr + p.sum()
                acc *= corr
                for jj in range(min(32, i - j0 + 1)):
                    acc += p[jj] * v[j0 + jj, hk].reshape(32, 4)
                m = m_new
            out[i, h] = (acc / l).reshape(d)
    s = np.einsum("ihc,jhc->hij", q, np.repeat(k, H // Hk, 1)) * d ** -0.5
    s = np.where(np.triu(np.ones((L, L), bool), 1)[None], -np.inf, s)
    p = np.exp(s - s.max(-1, keepdims=True))
    ref = np.einsum("hij,jhc->ihc", p / p.sum(-1, keepdims=True), np.repeat(v, H // Hk, 1))
    assert np.abs(out - ref).max() < 1e-12


def test_rope_tables_and_prefix_stripping():
    from wan2gp_b200.hyvideo.llm import rope_tables, strip_prefix
    for theta in (1e6, 5e5):
        c, s = rope_tables(33, theta)
        co, so = llm_oracle.rope_tables(33, theta)
        assert torch.equal(c, co) and torch.equal(s, so) and c.shape == (33, 64)
    cfg = synth.LLM_CONFIGS["qwen_tiny"]
    sd = synth.make_llm_state_dict(cfg, 0)
    for pre in ("", "model.", "model.language_model.", "language_model.model."):
        full = {pre + k: v for k, v in sd.items()}
        full.update({"lm_head.weight": sd["embed_tokens.weight"], "model.visual.blocks.0.attn.qkv.weight": torch.zeros(2, 2),
                     "visual.patch_embed.proj.weight": torch.zeros(1)})
        got = strip_prefix(full)
        assert set(got) == set(sd) and all(got[k] is sd[k] for k in sd), pre
    with pytest.raises(KeyError):
        strip_prefix({"foo.weight": torch.zeros(1)})


@pytest.mark.parametrize("name", ["qwen_tiny", "llama_tiny"])
def test_text_model_host_path(stub_abi, monkeypatch, name):  # noqa: F811
    """LlamaLikeTextModel with the stubbed C ABI: the transformers call surface the reference's TextEncoder.encode uses
    (text_encoder_1_5.py:470-482): model(input_ids=, attention_mask=, output_hidden_states=True).hidden_states[-3]."""
    import wan2gp_b200.hyvideo.llm as llm
    monkeypatch.setattr(llm, "_s", lambda: 0)
    cfg = synth.LLM_CONFIGS[name]
    sd = {"model.language_model." + k: v for k, v in synth.make_llm_state_dict(cfg, 0).items()}
    m = llm.LlamaLikeTextModel.from_state_dict(sd, cfg["num_heads"], cfg["num_kv_heads"], cfg["rms_eps"], cfg["rope_theta"], device="cpu")
    assert (m.vocab_size, m.hidden_size, m.intermediate_size, m.num_layers) == (cfg["vocab_size"], cfg["hidden_size"], cfg["intermediate_size"], cfg["num_layers"])
    assert (m.layers[0]["bqkv"] is not None) == cfg["qkv_bias"] and m.device.type == "cpu" and m.dtype == torch.bfloat16
    ids = torch.stack([synth.make_llm_inputs(cfg, 24, nv, s)[0] for s, nv in ((0, 9), (1, 24))])
    mask = torch.stack([synth.make_llm_inputs(cfg, 24, nv, s)[1] for s, nv in ((0, 9), (1, 24))])
    out = m(input_ids=ids, attention_mask=mask, output_hidden_states=True)
    assert len(out.hidden_states) == cfg["num_layers"] + 1 and tuple(out.hidden_states[-3].shape) == (2, 24, cfg["hidden_size"])
    assert float(out.hidden_states[-3][0, 9:].abs().max()) == 0.0                   # padded rows are returned as zeros
    assert stub_abi.count("b200_causal_gqa_attention") == 2 * cfg["num_layers"] and stub_abi.count("b200_rope_half") == 2 * cfg["num_layers"]
    assert m(input_ids=ids, attention_mask=mask).hidden_states is None
    with pytest.raises(NotImplementedError):                                        # left padding is not what the reference tokenizers produce
        m(input_ids=ids, attention_mask=mask.flip(1), output_hidden_states=True)
    with pytest.raises(ValueError):
        llm.LlamaLikeTextModel.from_state_dict(sd, cfg["num_heads"] * 2, cfg["num_kv_heads"], device="cpu")


@pytest.mark.parametrize("name", ["qwen_tiny", "llama_tiny"])
def test_oracle_matches_transformers_live(name):
    """The same pin as the fixture, live: transformers' own classes (installed in this image; on a box without them the test skips) on another
    seed, length and padding -- every hidden state on the valid rows."""
    transformers = pytest.importorskip("transformers")
    cfg = synth.LLM_CONFIGS[name]
    common = dict(vocab_size=cfg["vocab_size"], hidden_size=cfg["hidden_size"], intermediate_size=cfg["intermediate_size"],
                  num_hidden_layers=cfg["num_layers"], num_attention_heads=cfg["num_heads"], num_key_value_heads=cfg["num_kv_heads"],
                  rms_norm_eps=cfg["rms_eps"], rope_theta=cfg["rope_theta"], max_position_embeddings=4096, attn_implementation="eager",
                  tie_word_embeddings=False)
    try:
        if name.startswith("qwen"):
            from transformers.models.qwen2_5_vl.configuration_qwen2_5_vl import Qwen2_5_VLTextConfig
            from transformers.models.qwen2_5_vl.modeling_qwen2_5_vl import Qwen2_5_VLTextModel
            model = Qwen2_5_VLTextModel(Qwen2_5_VLTextConfig(rope_scaling={"type": "mrope", "mrope_section": [16, 24, 24]}, **common))
        else:
            model = transformers.LlamaModel(transformers.LlamaConfig(attention_bias=False, mlp_bias=False, head_dim=128, **common))
    except (ImportError, TypeError) as e:                       # another transformers version without these classes / arguments
        pytest.skip(f"transformers {transformers.__version__}: {e}")
    model = model.eval().float()
    sd = synth.make_llm_state_dict(cfg, seed=5)
    model.load_state_dict(sd, strict=True)
    ids, mask = synth.make_llm_inputs(cfg, 33, 19, seed=5)
    with torch.no_grad():
        out = model(input_ids=ids[None], attention_mask=mask[None], output_hidden_states=True)
    hs = llm_oracle.llm_hidden_states(sd, cfg, ids, 19)
    for i, (h, r) in enumerate(zip(hs, out.hidden_states)):
        assert float((h[:19] - r[0, :19]).norm() / r[0, :19].norm()) < 1e-6, i
