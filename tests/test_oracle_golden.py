"""CPU: the oracle restatement is pinned against outputs of the UNMODIFIED reference
(tests/golden/*.npz, made by oracle/gen_golden.py from /root/reference)."""
import pytest
import torch

from oracle import vae_oracle, wan_oracle
from tests.helpers import load_golden, rel_l2, vae_case, wan_case
from wan2gp_b200 import synth


@pytest.mark.parametrize("name", ["tiny", "tiny_i2v", "small"])
def test_wan_oracle_matches_reference(name):
    cfg, thw, sd, x, t, ctx, y = wan_case(name)
    g = load_golden("wan_" + name)
    cos, sin = wan_oracle.rope_tables(thw)
    assert torch.equal(cos[:64], g["cos"]) and torch.equal(sin[:64], g["sin"])        # W0 tables bit-exact
    # reference fp64 run == production RMSNorm semantics (no fp32 aliasing), tolerance 5e-6 rel-L2
    out = wan_oracle.wan_forward(sd, cfg, x, t, ctx, y=y)
    assert rel_l2(out, g["out64"]) < 5e-6
    # reference fp32 run, reproducing the model.py:165-172 aliasing artefact
    wan_oracle.FP32_ALIAS_QUIRK = True
    try:
        outq = wan_oracle.wan_forward(sd, cfg, x, t, ctx, y=y)
    finally:
        wan_oracle.FP32_ALIAS_QUIRK = False
    assert rel_l2(outq, g["out"]) < 5e-6
    # the bf16-emulating oracle (what the CUDA path is compared with) stays within 2e-3 of fp32
    oute = wan_oracle.wan_forward(sd, cfg, x, t, ctx, y=y, emulate_bf16=True)
    assert rel_l2(oute, g["out64"]) < 2e-3


def test_wan_oracle_p13b_matches_reference():
    """BASELINE config 1: Wan2.1 t2v 1.3B, latent [1,16,9,30,52], 30 blocks."""
    cfg, thw, sd, x, t, ctx, y = wan_case("p13b")
    g = load_golden("wan_p13b")
    out = wan_oracle.wan_forward(sd, cfg, x, t, ctx, y=y)
    assert rel_l2(out, g["out64"]) < 2e-5


@pytest.mark.parametrize("name", ["vae_tiny", "vae_small"])
def test_vae_oracle_matches_reference(name):
    cfg, sd, z = vae_case(name)
    g = load_golden(name)
    out = vae_oracle.vae_decode(sd, z, synth.VAE_MEAN, synth.VAE_STD, cfg)
    assert out.shape == g["out"][0].shape
    assert rel_l2(out, g["out"][0]) < 1e-5
    d = (vae_oracle.frames_to_uint8(out).int() - vae_oracle.frames_to_uint8(g["out"][0]).int()).abs()
    assert d.max() <= 1 and d.float().mean() < 1e-3


def test_cfg_and_euler():
    c, u = torch.randn(1, 16, 2, 4, 4), torch.randn(1, 16, 2, 4, 4)
    o = wan_oracle.cfg_combine(c, u, 4.0)
    assert torch.allclose(o, u + 4.0 * (c - u))
    o2 = wan_oracle.cfg_combine(c, u, 4.0, cfg_star=True, step_no=3)
    alpha = (c * u).sum() / (u.pow(2).sum() + 1e-8)
    assert torch.allclose(o2, alpha * u + 4.0 * (c - alpha * u), atol=1e-5)


@pytest.mark.parametrize("name", ["tiny", "small"])
def test_product_rope_tables_match_reference(name):
    """wan2gp_b200.wan.rope (product code) reproduces the reference's get_rotary_pos_embed tables bit-exactly."""
    from tests.helpers import WAN_CASES
    from wan2gp_b200.wan.rope import get_rotary_pos_embed
    _, thw, _ = WAN_CASES[name]
    g = load_golden("wan_" + name)
    cos, sin = get_rotary_pos_embed(thw)
    assert torch.equal(cos[:64], g["cos"]) and torch.equal(sin[:64], g["sin"])
    c2, s2 = wan_oracle.rope_tables(thw)
    assert torch.equal(cos, c2) and torch.equal(sin, s2)


def test_hy_oracle_matches_reference():
    """Hunyuan Video 1.5 DiT oracle vs the reference run (fp32 weights + the reference's own bf16 hard-casts)."""
    from oracle import hy_oracle
    from wan2gp_b200 import synth
    cfg, thw = synth.HY_CONFIGS["hy_tiny"], (3, 6, 10)
    sd = synth.make_hy_state_dict(cfg, 0)
    x, t, txt, tm, b5, bm = synth.make_hy_inputs(cfg, thw, seed=0)
    g = load_golden("hy_tiny")
    cos, sin = hy_oracle.rope_tables_hy(thw)
    assert torch.equal(cos[:64], g["cos"]) and torch.equal(sin[:64], g["sin"])
    out = hy_oracle.hy_forward(sd, cfg, x, t, txt, tm, b5, bm, ref_casts=True)
    assert rel_l2(out, g["out"]) < 2e-5
    assert rel_l2(hy_oracle.hy_forward(sd, cfg, x, t, txt, tm, b5, bm, emulate_bf16=True), g["out"]) < 5e-3


def test_hy10_oracle_matches_reference():
    """HunyuanVideo 1.0 family (double + single-stream blocks; mmgp's load-time Linear split restated in oracle/refshim.py)."""
    from oracle import hy_oracle
    from wan2gp_b200 import synth
    cfg, thw, seed = synth.HY_CONFIGS["hy10_tiny"], (2, 8, 12), 1
    sd = synth.make_hy_state_dict(cfg, seed)
    x, t, txt, tm, _, _ = synth.make_hy_inputs(cfg, thw, seed=seed)
    t2 = synth._normal((1, cfg["text_states_dim_2"]), 1.0, seed, "hy.txt2", "cpu")
    out = hy_oracle.hy_forward(sd, cfg, x, t, txt, tm, ref_casts=True, text_states_2=t2, guidance=torch.tensor([6000.0]))
    assert rel_l2(out, load_golden("hy10_tiny")["out"]) < 3e-5


@pytest.mark.parametrize("name,zshape,seed", [("hyvae_tiny", (8, 3, 4, 6), 0), ("hyvae_small", (16, 3, 2, 3), 1)])
def test_hyvae_oracle_matches_reference(name, zshape, seed):
    from oracle import hyvae_oracle
    from wan2gp_b200 import synth
    cfg = synth.HYVAE_CONFIGS[name]
    sd = synth.make_hyvae_state_dict(cfg, seed)
    z = synth._normal((1,) + zshape, 1.0, seed, "input.z", "cpu")[0]
    assert rel_l2(hyvae_oracle.hyvae_decode(sd, cfg, z), load_golden(name)["out"][0]) < 1e-5


def test_hy_i2v_oracle_matches_reference_fixture():
    """hunyuan_1_5_i2v: the reference HYVideoDiffusionTransformer with vision_projection='linear' and 200 image-encoder tokens
    (tests/golden/hy_tiny_i2v.npz, oracle/gen_golden.py) -- VisionProjection, cond-type embedding 2, tokens in front of the text stream."""
    from oracle import hy_oracle
    from tests.helpers import load_golden, rel_l2
    from wan2gp_b200 import synth
    cfg = synth.HY_CONFIGS["hy_tiny_i2v"]
    sd = synth.make_hy_state_dict(cfg, 2)
    x, t, txt, tm, b5, bm = synth.make_hy_inputs(cfg, (3, 6, 10), seed=2)
    vs = synth.make_hy_vision_states(cfg, seed=2)
    g = load_golden("hy_tiny_i2v")["out"]
    assert rel_l2(hy_oracle.hy_forward(sd, cfg, x, t, txt, tm, b5, bm, vision_states=vs), g) < 5e-5
    assert rel_l2(hy_oracle.hy_forward(sd, cfg, x, t, txt, tm, b5, bm), g) > 1e-2           # the vision tokens matter in this fixture
