"""umT5 text encoder: the oracle against the committed reference fixture, and the host-side pieces of
wan2gp_b200/wan/t5.py that need no GPU."""
import os

import numpy as np
import pytest
import torch

from oracle import refshim, t5_oracle
from wan2gp_b200 import synth

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _case(name="t5_small"):
    g = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    cfg = synth.T5_CONFIGS[name]
    sd = synth.make_t5_state_dict(cfg, int(g["seed"]))
    ids, mask = synth.make_t5_inputs(cfg, int(g["length"]), int(g["n_valid"]), int(g["seed"]))
    return cfg, sd, ids, mask, torch.from_numpy(g["out"])


def test_oracle_matches_reference_fixture():
    """tests/golden/t5_small.npz was produced by the UNMODIFIED reference T5Encoder (oracle/gen_golden.py t5_small)."""
    cfg, sd, ids, mask, ref = _case()
    out = t5_oracle.t5_encode(sd, cfg, ids, mask)
    assert float((out - ref).norm() / ref.norm()) < 1e-6


@pytest.mark.skipif(not refshim.reference_available(), reason="reference tree not present")
def test_oracle_matches_reference_module_live():
    cfg, sd, ids, mask, _ = _case()
    R = refshim.load_reference_t5()
    enc = R.T5Encoder(cfg["vocab_size"], cfg["dim"], cfg["dim_attn"], cfg["dim_ffn"], cfg["num_heads"], cfg["num_layers"], cfg["num_buckets"],
                      shared_pos=False).eval().float()
    enc.load_state_dict(sd)
    with torch.no_grad():
        ref = enc(ids[None], mask[None])[0]
    # a different padding length and an unmasked call as well
    out = t5_oracle.t5_encode(sd, cfg, ids, mask)
    assert float((out - ref).norm() / ref.norm()) < 1e-6
    with torch.no_grad():
        ref2 = enc(ids[None, :17])[0]
    assert float((t5_oracle.t5_encode(sd, cfg, ids[:17]) - ref2).norm() / ref2.norm()) < 1e-6


def test_relative_bias_table_is_the_expanded_bias():
    from wan2gp_b200.wan.t5 import relative_bias_table
    emb = synth._normal((32, 6), 0.5, 0, "pos", "cpu")
    for L in (1, 7, 40, 200):
        tab = relative_bias_table(emb, L, 32)                       # [heads, 2L-1]
        full = t5_oracle.position_bias(emb, L, L, 32)               # [heads, L, L]
        i, j = torch.meshgrid(torch.arange(L), torch.arange(L), indexing="ij")
        assert torch.equal(tab[:, j - i + L - 1], full)


def test_hf_names_map_to_reference_names():
    from wan2gp_b200.wan.t5 import hf_to_wan_names
    cfg = synth.T5_CONFIGS["t5_small"]
    sd = synth.make_t5_state_dict(cfg, 0)
    hf = {"shared.weight": sd["token_embedding.weight"], "encoder.final_layer_norm.weight": sd["norm.weight"]}
    for i in range(cfg["num_layers"]):
        b, h = f"blocks.{i}.", f"encoder.block.{i}.layer."
        hf[h + "0.layer_norm.weight"] = sd[b + "norm1.weight"]
        hf[h + "1.layer_norm.weight"] = sd[b + "norm2.weight"]
        for n in "qkvo":
            hf[h + f"0.SelfAttention.{n}.weight"] = sd[b + f"attn.{n}.weight"]
        hf[h + "0.SelfAttention.relative_attention_bias.weight"] = sd[b + "pos_embedding.embedding.weight"]
        hf[h + "1.DenseReluDense.wi_0.weight"] = sd[b + "ffn.gate.0.weight"]
        hf[h + "1.DenseReluDense.wi_1.weight"] = sd[b + "ffn.fc1.weight"]
        hf[h + "1.DenseReluDense.wo.weight"] = sd[b + "ffn.fc2.weight"]
    back = hf_to_wan_names(hf)
    assert set(back) == set(sd) and all(back[k] is sd[k] for k in sd)
    assert hf_to_wan_names(sd) is sd


def test_clean_prompt_is_the_reference_whitespace_clean():
    from wan2gp_b200.wan.t5 import clean_prompt
    assert clean_prompt("  a &amp;amp; b \n\t c  ") == "a & b c"
    assert clean_prompt("plain prompt") == "plain prompt"
