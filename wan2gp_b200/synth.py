"""Seeded synthetic weights and inputs (there are no checkpoints offline).

State-dict names and shapes follow the reference modules exactly (SURVEY.md
Appendix B; models/wan/modules/model.py:1131-1143, 530-551, 847-850 and
models/wan/modules/vae.py:430-484) so the same dict loads into the reference
`WanModel` / `WanVAE_` (oracle side) and into this package's modules.

Initialisation is the reference's `init_weights` (model.py:2128-2150) with the
"tempered" overrides of SURVEY.md section 7/H1: with the stock init the QK logits
have std ~ sqrt(128) and softmax is near-argmax, which makes any bf16-vs-fp32
comparison chaotic; `norm_q/k.weight ~ 0.3` keeps attention in a realistic regime.
Every tensor has its own generator seeded from (seed, crc32(name)) so a single
tensor can be regenerated anywhere (CPU values are bit-reproducible across hosts).
"""
import math
import zlib

import torch

WAN_CONFIGS = {
    # models/wan/configs/t2v_1.3B.json
    "t2v_1.3B": dict(model_type="t2v", dim=1536, ffn_dim=8960, freq_dim=256, in_dim=16, out_dim=16,
                     num_heads=12, num_layers=30, text_len=512, text_dim=4096, eps=1e-6),
    # models/wan/configs/t2v_2_2.json (Wan2.2 14B, high- and low-noise experts share it)
    "t2v_2_2": dict(model_type="t2v", dim=5120, ffn_dim=13824, freq_dim=256, in_dim=16, out_dim=16,
                    num_heads=40, num_layers=40, text_len=512, text_dim=4096, eps=1e-6),
    # models/wan/configs/i2v_2_2.json
    "i2v_2_2": dict(model_type="i2v2_2", dim=5120, ffn_dim=13824, freq_dim=256, in_dim=36, out_dim=16,
                    num_heads=40, num_layers=40, text_len=512, text_dim=4096, eps=1e-6),
    # reduced configs for fast parity tests (head_dim stays 128 as in every Wan model)
    "tiny": dict(model_type="t2v", dim=256, ffn_dim=768, freq_dim=256, in_dim=16, out_dim=16,
                 num_heads=2, num_layers=2, text_len=64, text_dim=128, eps=1e-6),
    "tiny_i2v": dict(model_type="i2v2_2", dim=256, ffn_dim=768, freq_dim=256, in_dim=36, out_dim=16,
                     num_heads=2, num_layers=2, text_len=64, text_dim=128, eps=1e-6),
    "small": dict(model_type="t2v", dim=512, ffn_dim=1536, freq_dim=256, in_dim=16, out_dim=16,
                  num_heads=4, num_layers=3, text_len=512, text_dim=512, eps=1e-6),
}


def _gen(seed, name, device):
    g = torch.Generator(device=device)
    g.manual_seed((int(seed) * 1000003 + zlib.crc32(name.encode())) % (2 ** 63 - 1))
    return g


def _normal(shape, std, seed, name, device, mean=0.0):
    t = torch.empty(shape, dtype=torch.float32, device=device)
    t.normal_(mean, std, generator=_gen(seed, name, device))
    return t


def _xavier(shape, seed, name, device):
    fan_out, fan_in = shape[0], math.prod(shape[1:])
    a = math.sqrt(6.0 / (fan_in + fan_out))
    t = torch.empty(shape, dtype=torch.float32, device=device)
    t.uniform_(-a, a, generator=_gen(seed, name, device))
    return t


def wan_param_shapes(cfg):
    """name -> shape for the t2v / i2v2_2 WanModel."""
    D, F, Cin, Td, Fd = cfg["dim"], cfg["ffn_dim"], cfg["in_dim"], cfg["text_dim"], cfg["freq_dim"]
    s = {
        "patch_embedding.weight": (D, Cin, 1, 2, 2), "patch_embedding.bias": (D,),
        "text_embedding.0.weight": (D, Td), "text_embedding.0.bias": (D,),
        "text_embedding.2.weight": (D, D), "text_embedding.2.bias": (D,),
        "time_embedding.0.weight": (D, Fd), "time_embedding.0.bias": (D,),
        "time_embedding.2.weight": (D, D), "time_embedding.2.bias": (D,),
        "time_projection.1.weight": (6 * D, D), "time_projection.1.bias": (6 * D,),
        "head.modulation": (1, 2, D),
        "head.head.weight": (4 * cfg["out_dim"], D), "head.head.bias": (4 * cfg["out_dim"],),
    }
    for i in range(cfg["num_layers"]):
        p = f"blocks.{i}."
        s[p + "modulation"] = (1, 6, D)
        for a in ("self_attn", "cross_attn"):
            for l in "qkvo":
                s[p + f"{a}.{l}.weight"] = (D, D)
                s[p + f"{a}.{l}.bias"] = (D,)
            s[p + f"{a}.norm_q.weight"] = (D,)
            s[p + f"{a}.norm_k.weight"] = (D,)
        s[p + "norm3.weight"] = (D,)
        s[p + "norm3.bias"] = (D,)
        s[p + "ffn.0.weight"] = (F, D)
        s[p + "ffn.0.bias"] = (F,)
        s[p + "ffn.2.weight"] = (D, F)
        s[p + "ffn.2.bias"] = (D,)
    return s


def make_wan_tensor(name, shape, cfg, seed=0, device="cpu"):
    D = cfg["dim"]
    if name.endswith("modulation"):
        return _normal(shape, 1.0 / math.sqrt(D), seed, name, device)
    if "norm_q" in name or "norm_k" in name:
        return 0.3 * (1.0 + _normal(shape, 0.1, seed, name, device))
    if name.endswith("norm3.weight"):
        return 1.0 + _normal(shape, 0.1, seed, name, device)
    if name.endswith("bias"):
        return _normal(shape, 0.02, seed, name, device)
    if name.startswith(("text_embedding", "time_embedding")) or name == "head.head.weight":
        return _normal(shape, 0.02, seed, name, device)
    return _xavier(shape, seed, name, device)  # Linear / patch-embed weights


def make_wan_state_dict(cfg, seed=0, device="cpu", dtype=torch.float32):
    return {n: make_wan_tensor(n, s, cfg, seed, device).to(dtype) for n, s in wan_param_shapes(cfg).items()}


# --------------------------------------------------------------------------- VAE

VAE_CFG = dict(dim=96, z_dim=16, dim_mult=[1, 2, 4, 4], num_res_blocks=2)  # vae.py:911-918
VAE_CFG_TINY = dict(dim=32, z_dim=16, dim_mult=[1, 2, 4, 4], num_res_blocks=2)

VAE_MEAN = [-0.7571, -0.7089, -0.9113, 0.1075, -0.1745, 0.9653, -0.1517, 1.5508,
            0.4134, -0.0715, 0.5517, -0.3632, -0.1922, -0.9497, 0.2503, -0.2921]  # vae.py:948-951
VAE_STD = [2.8184, 1.4541, 2.3275, 2.6558, 1.2196, 1.7708, 2.6052, 2.0743,
           3.2687, 2.1526, 2.8652, 1.5579, 1.6382, 1.1253, 2.8251, 1.9160]       # vae.py:952-955


def vae_decoder_layout(cfg=VAE_CFG):
    """Structural description of Decoder3d (vae.py:430-484): list of
    ("res", cin, cout) | ("attn", c) | ("up3d", c) | ("up2d", c) for `upsamples`,
    plus the channel count of conv1/middle and of the head."""
    dim, mult = cfg["dim"], cfg["dim_mult"]
    dims = [dim * u for u in [mult[-1]] + mult[::-1]]
    temporal = [False, True, True][::-1]  # temperal_downsample reversed (vae.py:573)
    ups = []
    for i, (cin, cout) in enumerate(zip(dims[:-1], dims[1:])):
        if i in (1, 2, 3):
            cin = cin // 2
        for _ in range(cfg["num_res_blocks"] + 1):
            ups.append(("res", cin, cout))
            cin = cout
        if i != len(mult) - 1:
            ups.append(("up3d" if temporal[i] else "up2d", cout))
    return dims[0], ups, dims[-1]


def vae_param_shapes(cfg=VAE_CFG):
    z = cfg["z_dim"]
    c0, ups, c_out = vae_decoder_layout(cfg)
    s = {"conv2.weight": (z, z, 1, 1, 1), "conv2.bias": (z,),
         "decoder.conv1.weight": (c0, z, 3, 3, 3), "decoder.conv1.bias": (c0,)}

    def res(p, ci, co):
        s[p + "residual.0.gamma"] = (ci, 1, 1, 1)
        s[p + "residual.2.weight"] = (co, ci, 3, 3, 3)
        s[p + "residual.2.bias"] = (co,)
        s[p + "residual.3.gamma"] = (co, 1, 1, 1)
        s[p + "residual.6.weight"] = (co, co, 3, 3, 3)
        s[p + "residual.6.bias"] = (co,)
        if ci != co:
            s[p + "shortcut.weight"] = (co, ci, 1, 1, 1)
            s[p + "shortcut.bias"] = (co,)

    res("decoder.middle.0.", c0, c0)
    s["decoder.middle.1.norm.gamma"] = (c0, 1, 1)
    s["decoder.middle.1.to_qkv.weight"] = (3 * c0, c0, 1, 1)
    s["decoder.middle.1.to_qkv.bias"] = (3 * c0,)
    s["decoder.middle.1.proj.weight"] = (c0, c0, 1, 1)
    s["decoder.middle.1.proj.bias"] = (c0,)
    res("decoder.middle.2.", c0, c0)
    for j, u in enumerate(ups):
        p = f"decoder.upsamples.{j}."
        if u[0] == "res":
            res(p, u[1], u[2])
        else:
            c = u[1]
            s[p + "resample.1.weight"] = (c // 2, c, 3, 3)
            s[p + "resample.1.bias"] = (c // 2,)
            if u[0] == "up3d":
                s[p + "time_conv.weight"] = (2 * c, c, 3, 1, 1)
                s[p + "time_conv.bias"] = (2 * c,)
    s["decoder.head.0.gamma"] = (c_out, 1, 1, 1)
    s["decoder.head.2.weight"] = (3, c_out, 3, 3, 3)
    s["decoder.head.2.bias"] = (3,)
    return s


def vae_encoder_layout(cfg=VAE_CFG):
    """Encoder3d (vae.py:318-372): list of ("res", cin, cout) | ("down2d", c) | ("down3d", c) for `downsamples`, plus the
    channel count of conv1 and of the middle/head; temperal_downsample = [False, True, True] (vae.py:911-918)."""
    dim, mult = cfg["dim"], cfg["dim_mult"]
    dims = [dim * u for u in [1] + mult]
    temporal = [False, True, True]
    downs = []
    for i, (cin, cout) in enumerate(zip(dims[:-1], dims[1:])):
        for _ in range(cfg["num_res_blocks"]):
            downs.append(("res", cin, cout))
            cin = cout
        if i != len(mult) - 1:
            downs.append(("down3d" if temporal[i] else "down2d", cout))
    return dims[0], downs, dims[-1]


def vae_encoder_param_shapes(cfg=VAE_CFG):
    """encoder.* and conv1.* of WanVAE_ (vae.py:318-372, 582)."""
    z = cfg["z_dim"]
    c0, downs, c_mid = vae_encoder_layout(cfg)
    s = {"conv1.weight": (2 * z, 2 * z, 1, 1, 1), "conv1.bias": (2 * z,),
         "encoder.conv1.weight": (c0, 3, 3, 3, 3), "encoder.conv1.bias": (c0,)}

    def res(p, ci, co):
        s[p + "residual.0.gamma"] = (ci, 1, 1, 1)
        s[p + "residual.2.weight"] = (co, ci, 3, 3, 3)
        s[p + "residual.2.bias"] = (co,)
        s[p + "residual.3.gamma"] = (co, 1, 1, 1)
        s[p + "residual.6.weight"] = (co, co, 3, 3, 3)
        s[p + "residual.6.bias"] = (co,)
        if ci != co:
            s[p + "shortcut.weight"] = (co, ci, 1, 1, 1)
            s[p + "shortcut.bias"] = (co,)
    for j, u in enumerate(downs):
        p = f"encoder.downsamples.{j}."
        if u[0] == "res":
            res(p, u[1], u[2])
        else:
            c = u[1]
            s[p + "resample.1.weight"] = (c, c, 3, 3)
            s[p + "resample.1.bias"] = (c,)
            if u[0] == "down3d":
                s[p + "time_conv.weight"] = (c, c, 3, 1, 1)
                s[p + "time_conv.bias"] = (c,)
    res("encoder.middle.0.", c_mid, c_mid)
    s["encoder.middle.1.norm.gamma"] = (c_mid, 1, 1)
    s["encoder.middle.1.to_qkv.weight"] = (3 * c_mid, c_mid, 1, 1)
    s["encoder.middle.1.to_qkv.bias"] = (3 * c_mid,)
    s["encoder.middle.1.proj.weight"] = (c_mid, c_mid, 1, 1)
    s["encoder.middle.1.proj.bias"] = (c_mid,)
    res("encoder.middle.2.", c_mid, c_mid)
    s["encoder.head.0.gamma"] = (c_mid, 1, 1, 1)
    s["encoder.head.2.weight"] = (2 * z, c_mid, 3, 3, 3)
    s["encoder.head.2.bias"] = (2 * z,)
    return s


def make_vae_tensor(name, shape, seed=0, device="cpu"):
    if name.endswith("gamma"):
        return 1.0 + _normal(shape, 0.1, seed, name, device)
    if name.endswith("bias"):
        return _normal(shape, 0.02, seed, name, device)
    # PyTorch default conv init (kaiming_uniform a=sqrt(5)) == U(-1/sqrt(fan_in), 1/sqrt(fan_in)),
    # scaled x1.7 so activations keep O(1) magnitude through 30 conv layers of random weights
    fan_in = math.prod(shape[1:])
    a = 1.7 / math.sqrt(fan_in)
    t = torch.empty(shape, dtype=torch.float32, device=device)
    t.uniform_(-a, a, generator=_gen(seed, name, device))
    return t


def make_vae_state_dict(cfg=VAE_CFG, seed=0, device="cpu", dtype=torch.float32, encoder=False):
    shapes = dict(vae_param_shapes(cfg))
    if encoder:
        shapes.update(vae_encoder_param_shapes(cfg))
    return {n: make_vae_tensor(n, s, seed, device).to(dtype) for n, s in shapes.items()}


# ------------------------------------------------------------------------ inputs

def make_wan_inputs(cfg, latent_shape, seed=0, batch=1):
    """latent x~N(0,1) [B,16,T,H,W]; context ~N(0,1) [1,text_len,text_dim]; t=500;
    i2v: y [20,T,H,W] with first 4 channels in {0,1}."""
    T, H, W = latent_shape
    x = _normal((batch, 16, T, H, W), 1.0, seed, "input.x", "cpu")
    ctx = _normal((1, cfg["text_len"], cfg["text_dim"]), 1.0, seed, "input.context", "cpu")
    t = torch.tensor([500.0])
    y = None
    if cfg["in_dim"] > 16:
        y = _normal((cfg["in_dim"] - 16, T, H, W), 1.0, seed, "input.y", "cpu")
        y[:4] = (y[:4] > 0).float()
    return x, t, ctx, y


# --------------------------------------------------------------------------- Hunyuan Video 1.5 (double-stream DiT)

HY_CONFIGS = {
    # models/hyvideo/modules/models.py:1345-1363 'HYVideo-1_5' + hunyuan.py:230-233 (in 65 / out 32 channels)
    "HYVideo-1_5": dict(hidden_size=2048, heads_num=16, mlp_width_ratio=4, mm_double_blocks_depth=54, in_channels=65,
                        out_channels=32, text_states_dim=3584, patch_size=[1, 1, 1], rope_dim_list=[16, 56, 56]),
    "hy_tiny": dict(hidden_size=256, heads_num=2, mlp_width_ratio=4, mm_double_blocks_depth=2, in_channels=65,
                    out_channels=32, text_states_dim=96, patch_size=[1, 1, 1], rope_dim_list=[16, 56, 56]),
}
HY_CONFIGS["hy10_tiny"] = dict(hidden_size=256, heads_num=2, mlp_width_ratio=4, mm_double_blocks_depth=1, mm_single_blocks_depth=2,
                               in_channels=16, out_channels=16, text_states_dim=96, text_states_dim_2=64, patch_size=[1, 2, 2],
                               rope_dim_list=[16, 56, 56], guidance_embed=True, family="1.0")
# models/hyvideo/modules/models.py:1287-1295 'HYVideo-T/2-cfgdistill' (HunyuanVideo 1.0: 20 double + 40 single blocks)
HY_CONFIGS["HYVideo-T/2-cfgdistill"] = dict(hidden_size=3072, heads_num=24, mlp_width_ratio=4, mm_double_blocks_depth=20,
                                            mm_single_blocks_depth=40, in_channels=16, out_channels=16, text_states_dim=4096,
                                            text_states_dim_2=768, patch_size=[1, 2, 2], rope_dim_list=[16, 56, 56],
                                            guidance_embed=True, family="1.0")
HY_CONFIGS["hy_tiny_i2v"] = dict(HY_CONFIGS["hy_tiny"], vision_states_dim=64)      # + the image-encoder token projection of hunyuan_1_5_i2v
HY_BYT5_DIMS = (1472, 2048, 2048)      # ByT5Mapper(in_dim, hidden_dim, out_dim) constants, models.py:647-653


def hy_param_shapes(cfg):
    """name -> shape of HYVideoDiffusionTransformer('HYVideo-1_5' family: pre-split qkv, byT5 mapper, cond-type embedding,
    token refiner depth 2), SURVEY.md Appendix B.  vision_in.* is not listed (only used with vision_states)."""
    D, Td, Cin, Co = cfg["hidden_size"], cfg["text_states_dim"], cfg["in_channels"], cfg["out_channels"]
    F = int(D * cfg["mlp_width_ratio"])
    bi, bh, bo = HY_BYT5_DIMS
    pp = cfg["patch_size"][0] * cfg["patch_size"][1] * cfg["patch_size"][2]
    v10 = cfg.get("family") == "1.0"          # HunyuanVideo 1.0: fused qkv, single-stream blocks, pooled-text + guidance vectors
    s = {}

    def lin(name, o, i):
        s[name + ".weight"] = (o, i)
        s[name + ".bias"] = (o,)
    if not v10:
        s["byt5_in.layernorm.weight"] = s["byt5_in.layernorm.bias"] = (bi,)
        lin("byt5_in.fc1", bh, bi), lin("byt5_in.fc2", bo, bh), lin("byt5_in.fc3", D, bo)
    else:
        lin("vector_in.in_layer", D, cfg["text_states_dim_2"]), lin("vector_in.out_layer", D, D)
        if cfg.get("guidance_embed"):
            lin("guidance_in.mlp.0", D, 256), lin("guidance_in.mlp.2", D, D)
    s["img_in.proj.weight"] = (D, Cin, *cfg["patch_size"])
    s["img_in.proj.bias"] = (D,)
    lin("txt_in.input_embedder", D, Td)
    lin("txt_in.t_embedder.mlp.0", D, 256), lin("txt_in.t_embedder.mlp.2", D, D)
    lin("txt_in.c_embedder.linear_1", D, Td), lin("txt_in.c_embedder.linear_2", D, D)
    for j in range(2):
        p = f"txt_in.individual_token_refiner.blocks.{j}."
        s[p + "norm1.weight"] = s[p + "norm1.bias"] = s[p + "norm2.weight"] = s[p + "norm2.bias"] = (D,)
        lin(p + "self_attn_qkv", 3 * D, D), lin(p + "self_attn_proj", D, D)
        lin(p + "mlp.fc1", F, D), lin(p + "mlp.fc2", D, F), lin(p + "adaLN_modulation.1", 2 * D, D)
    lin("time_in.mlp.0", D, 256), lin("time_in.mlp.2", D, D)
    for i in range(cfg["mm_double_blocks_depth"]):
        for st in ("img", "txt"):
            p = f"double_blocks.{i}.{st}_"
            lin(p + "mod.linear", 6 * D, D)
            if v10:
                lin(p + "attn_qkv", 3 * D, D)
            else:
                for l in "qkv":
                    lin(p + "attn_" + l, D, D)
            s[p + "attn_q_norm.weight"] = s[p + "attn_k_norm.weight"] = (D // cfg["heads_num"],)
            lin(p + "attn_proj", D, D), lin(p + "mlp.fc1", F, D), lin(p + "mlp.fc2", D, F)
    for i in range(cfg.get("mm_single_blocks_depth", 0)):
        p = f"single_blocks.{i}."
        lin(p + "linear1", 3 * D + F, D), lin(p + "linear2", D, D + F), lin(p + "modulation.linear", 3 * D, D)
        s[p + "q_norm.weight"] = s[p + "k_norm.weight"] = (D // cfg["heads_num"],)
    lin("final_layer.linear", pp * Co, D), lin("final_layer.adaLN_modulation.1", 2 * D, D)
    if not v10:
        s["cond_type_embedding.weight"] = (3, D)
    if cfg.get("vision_states_dim"):                            # VisionProjection (embed_layers.py:62-77), Hunyuan 1.5 i2v
        Dv = cfg["vision_states_dim"]
        s["vision_in.proj.0.weight"] = s["vision_in.proj.0.bias"] = (Dv,)
        lin("vision_in.proj.1", Dv, Dv), lin("vision_in.proj.3", D, Dv)
        s["vision_in.proj.4.weight"] = s["vision_in.proj.4.bias"] = (D,)
    return s


def make_hy_tensor(name, shape, seed=0, device="cpu"):
    if name.endswith("norm.weight") or name.endswith("norm1.weight") or name.endswith("norm2.weight") or name.endswith("layernorm.weight") \
            or name in ("vision_in.proj.0.weight", "vision_in.proj.4.weight"):
        return 1.0 + _normal(shape, 0.1, seed, name, device)
    if name.endswith("bias"):
        return _normal(shape, 0.02, seed, name, device)
    if "mod.linear" in name or "modulation.linear" in name or "adaLN_modulation" in name or name.startswith(("final_layer.linear", "cond_type_embedding")):
        return _normal(shape, 0.02, seed, name, device)          # zero-initialised in the reference; randomised so they matter
    fan_in = math.prod(shape[1:])
    return _normal(shape, 1.0 / math.sqrt(fan_in), seed, name, device)


def make_hy_state_dict(cfg, seed=0, device="cpu", dtype=torch.float32):
    return {n: make_hy_tensor(n, s, seed, device).to(dtype) for n, s in hy_param_shapes(cfg).items()}


def make_hy_inputs(cfg, latent_thw, n_txt=24, n_txt_valid=17, n_byt5=12, n_byt5_valid=5, seed=0):
    """x [1,Cin,T,H,W], t=[500], LLM states [1,n_txt,text_dim] + mask (valid prefix), byT5 states [1,n_byt5,1472] + mask."""
    T, H, W = latent_thw
    x = _normal((1, cfg["in_channels"], T, H, W), 1.0, seed, "hy.x", "cpu")
    txt = _normal((1, n_txt, cfg["text_states_dim"]), 1.0, seed, "hy.txt", "cpu")
    byt5 = _normal((1, n_byt5, HY_BYT5_DIMS[0]), 1.0, seed, "hy.byt5", "cpu")
    tm = torch.zeros(1, n_txt, dtype=torch.long)
    tm[:, :n_txt_valid] = 1
    bm = torch.zeros(1, n_byt5, dtype=torch.long)
    bm[:, :n_byt5_valid] = 1
    return x, torch.tensor([500.0]), txt, tm, byt5, bm


def make_hy_vision_states(cfg, n_tokens=200, seed=0):
    """SigLIP-like image-encoder states [1, n_tokens, vision_states_dim] (729 x 1152 in production: hunyuan.py:305-307)."""
    return _normal((1, n_tokens, cfg["vision_states_dim"]), 1.0, seed, "hy.vision", "cpu")


# --------------------------------------------------------------------------- Hunyuan Video 1.5 VAE decoder (AutoencoderKLConv3D)

HYVAE_CONFIGS = {
    # hunyuan_video_1_5_VAE.json is a DOWNLOAD (hunyuan.py:329-336), not in the reference tree: these are the upstream values
    # (SURVEY.md section 8c), block_out_channels given in DECODER order
    "hyvae15": dict(z_channels=32, out_channels=3, block_out_channels=[1024, 1024, 512, 256, 128], num_res_blocks=2,
                    ffactor_spatial=16, ffactor_temporal=4),
    "hyvae_tiny": dict(z_channels=8, out_channels=3, block_out_channels=[64, 64, 32], num_res_blocks=1, ffactor_spatial=4,
                       ffactor_temporal=2),
    "hyvae_small": dict(z_channels=16, out_channels=3, block_out_channels=[128, 128, 64, 64, 32], num_res_blocks=2,
                        ffactor_spatial=16, ffactor_temporal=4),
}


def hyvae_layout(cfg):
    """Per level: (list of (cin, cout) resnet blocks, upsample (cin, cout, temporal) or None) -- Decoder.__init__
    (hunyuanvideo_15_vae.py:432-484)."""
    boc = cfg["block_out_channels"]
    n_sp, n_t = int(math.log2(cfg["ffactor_spatial"])), int(math.log2(cfg["ffactor_temporal"]))
    levels, cin = [], boc[0]
    for i, ch in enumerate(boc):
        blocks = []
        for _ in range(cfg["num_res_blocks"] + 1):
            blocks.append((cin, ch))
            cin = ch
        up = None
        if i < n_sp or i < n_t:
            up = (cin, boc[i + 1], i < n_t)
            cin = boc[i + 1]
        levels.append((blocks, up))
    return levels, cin


def hyvae_param_shapes(cfg):
    s = {}

    def conv(name, co, ci, k):
        s[name + ".weight"] = (co, ci, k, k, k)
        s[name + ".bias"] = (co,)

    def res(p, ci, co):
        s[p + "norm1.gamma"] = (ci, 1, 1, 1)
        conv(p + "conv1.conv", co, ci, 3)
        s[p + "norm2.gamma"] = (co, 1, 1, 1)
        conv(p + "conv2.conv", co, co, 3)
        if ci != co:
            conv(p + "nin_shortcut", co, ci, 1)
    c0 = cfg["block_out_channels"][0]
    conv("conv_in.conv", c0, cfg["z_channels"], 3)
    res("mid.block_1.", c0, c0)
    s["mid.attn_1.norm.gamma"] = (c0, 1, 1, 1)
    for n in ("q", "k", "v", "proj_out"):
        conv("mid.attn_1." + n, c0, c0, 1)
    res("mid.block_2.", c0, c0)
    levels, c_last = hyvae_layout(cfg)
    for i, (blocks, up) in enumerate(levels):
        for j, (ci, co) in enumerate(blocks):
            res(f"up.{i}.block.{j}.", ci, co)
        if up is not None:
            conv(f"up.{i}.upsample.conv.conv", up[1] * (8 if up[2] else 4), up[0], 3)
    s["norm_out.gamma"] = (c_last, 1, 1, 1)
    conv("conv_out.conv", cfg["out_channels"], c_last, 3)
    return s


def hyvae_encoder_layout(cfg):
    """Encoder.__init__ (hunyuanvideo_15_vae.py:345-393): per level (list of (cin, cout) resnets, downsample (cin, cout, temporal) or
    None), plus the channel count of the middle block.  HYVAE_CONFIGS hold block_out_channels in DECODER order."""
    boc = list(reversed(cfg["block_out_channels"]))
    n_sp, n_t0 = math.log2(cfg["ffactor_spatial"]), math.log2(cfg["ffactor_spatial"] // cfg["ffactor_temporal"])
    levels, cin = [], boc[0]
    for i, ch in enumerate(boc):
        blocks = []
        for _ in range(cfg["num_res_blocks"]):
            blocks.append((cin, ch))
            cin = ch
        down = None
        if i < n_sp:
            down = (cin, boc[i + 1], i >= n_t0)
            cin = boc[i + 1]
        levels.append((blocks, down))
    return levels, cin


def hyvae_encoder_param_shapes(cfg, in_channels=3):
    s = {}

    def conv(name, co, ci, k):
        s[name + ".weight"] = (co, ci, k, k, k)
        s[name + ".bias"] = (co,)

    def res(p, ci, co):
        s[p + "norm1.gamma"] = (ci, 1, 1, 1)
        conv(p + "conv1.conv", co, ci, 3)
        s[p + "norm2.gamma"] = (co, 1, 1, 1)
        conv(p + "conv2.conv", co, co, 3)
        if ci != co:
            conv(p + "nin_shortcut", co, ci, 1)
    levels, c_mid = hyvae_encoder_layout(cfg)
    conv("conv_in.conv", levels[0][0][0][0], in_channels, 3)
    for i, (blocks, down) in enumerate(levels):
        for j, (ci, co) in enumerate(blocks):
            res(f"down.{i}.block.{j}.", ci, co)
        if down is not None:
            conv(f"down.{i}.downsample.conv.conv", down[1] // (8 if down[2] else 4), down[0], 3)
    res("mid.block_1.", c_mid, c_mid)
    s["mid.attn_1.norm.gamma"] = (c_mid, 1, 1, 1)
    for n in ("q", "k", "v", "proj_out"):
        conv("mid.attn_1." + n, c_mid, c_mid, 1)
    res("mid.block_2.", c_mid, c_mid)
    s["norm_out.gamma"] = (c_mid, 1, 1, 1)
    conv("conv_out.conv", 2 * cfg["z_channels"], c_mid, 3)
    return s


def make_hyvae_state_dict(cfg, seed=0, device="cpu", dtype=torch.float32, encoder=False):
    shapes = hyvae_encoder_param_shapes(cfg) if encoder else hyvae_param_shapes(cfg)
    return {n: make_vae_tensor(("enc." if encoder else "") + n, s, seed, device).to(dtype) for n, s in shapes.items()}


# --------------------------------------------------------------------------- HunyuanVideo 1.0 VAE decoder (AutoencoderKLCausal3D)

HYVAE10_CONFIGS = {
    # "884-16c-hy" (models/hyvideo/vae/__init__.py:8; config.json is a DOWNLOAD): upstream values, block_out_channels in
    # ENCODER order as in the reference config
    "hyvae10": dict(latent_channels=16, out_channels=3, block_out_channels=[128, 256, 512, 512], layers_per_block=2,
                    norm_num_groups=32, time_compression_ratio=4, spatial_compression_ratio=8, mid_block_causal_attn=True),
    "hyvae10_tiny": dict(latent_channels=8, out_channels=3, block_out_channels=[32, 64, 64, 64], layers_per_block=1,
                         norm_num_groups=8, time_compression_ratio=4, spatial_compression_ratio=8, mid_block_causal_attn=True),
    "hyvae10_small": dict(latent_channels=16, out_channels=3, block_out_channels=[64, 128, 256, 256], layers_per_block=2,
                          norm_num_groups=32, time_compression_ratio=4, spatial_compression_ratio=8, mid_block_causal_attn=True),
}


def hyvae10_layout(cfg):
    """Per up block: (list of (cin, cout) resnets, (up_t, up_s) or None) -- DecoderCausal3D.__init__ (vae/vae.py:248-286),
    time_compression_ratio 4 rule."""
    boc = list(reversed(cfg["block_out_channels"]))
    assert cfg["time_compression_ratio"] == 4
    n_sp, n_t = int(math.log2(cfg["spatial_compression_ratio"])), int(math.log2(cfg["time_compression_ratio"]))
    blocks, cin = [], boc[0]
    for i, ch in enumerate(boc):
        res = []
        for _ in range(cfg["layers_per_block"] + 1):
            res.append((cin, ch))
            cin = ch
        sp, tm = i < n_sp, (i >= len(boc) - 1 - n_t and i != len(boc) - 1)
        blocks.append((res, (tm, sp) if (sp or tm) else None))
    return blocks, cin


def hyvae10_param_shapes(cfg):
    """State-dict names of AutoencoderKLCausal3D's decode half: post_quant_conv + decoder.* (vae/vae.py, unet_causal_3d_blocks.py)."""
    s = {}

    def conv(name, co, ci, k):
        s[name + ".weight"] = (co, ci, k, k, k)
        s[name + ".bias"] = (co,)

    def norm(name, c):
        s[name + ".weight"] = s[name + ".bias"] = (c,)

    def res(p, ci, co):
        norm(p + "norm1", ci), conv(p + "conv1.conv", co, ci, 3), norm(p + "norm2", co), conv(p + "conv2.conv", co, co, 3)
        if ci != co:
            conv(p + "conv_shortcut.conv", co, ci, 1)
    zc, c0 = cfg["latent_channels"], cfg["block_out_channels"][-1]
    conv("post_quant_conv", zc, zc, 1)
    conv("decoder.conv_in.conv", c0, zc, 3)
    res("decoder.mid_block.resnets.0.", c0, c0), res("decoder.mid_block.resnets.1.", c0, c0)
    a = "decoder.mid_block.attentions.0."
    norm(a + "group_norm", c0)
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        s[a + n + ".weight"], s[a + n + ".bias"] = (c0, c0), (c0,)
    blocks, c_last = hyvae10_layout(cfg)
    for i, (rs, up) in enumerate(blocks):
        for j, (ci, co) in enumerate(rs):
            res(f"decoder.up_blocks.{i}.resnets.{j}.", ci, co)
        if up is not None:
            conv(f"decoder.up_blocks.{i}.upsamplers.0.conv.conv", rs[-1][1], rs[-1][1], 3)
    norm("decoder.conv_norm_out", c_last)
    conv("decoder.conv_out.conv", cfg["out_channels"], c_last, 3)
    return s


def hyvae10_encoder_layout(cfg):
    """Per down block: (list of (cin, cout) resnets, (stride_t, stride_s) or None) -- EncoderCausal3D.__init__ (vae/vae.py:75-118),
    time_compression_ratio 4 rule."""
    boc = list(cfg["block_out_channels"])
    assert cfg["time_compression_ratio"] == 4
    n_sp, n_t = int(math.log2(cfg["spatial_compression_ratio"])), int(math.log2(cfg["time_compression_ratio"]))
    blocks, cin = [], boc[0]
    for i, ch in enumerate(boc):
        res = []
        for _ in range(cfg["layers_per_block"]):
            res.append((cin, ch))
            cin = ch
        sp, tm = i < n_sp, (i >= len(boc) - 1 - n_t and i != len(boc) - 1)
        blocks.append((res, (tm, sp) if (sp or tm) else None))
    return blocks, cin


def hyvae10_encoder_param_shapes(cfg, in_channels=3):
    """State-dict names of AutoencoderKLCausal3D's encode half: encoder.* + quant_conv (vae/vae.py:48-184, autoencoder_kl_causal_3d.py:243)."""
    s = {}

    def conv(name, co, ci, k):
        s[name + ".weight"] = (co, ci, k, k, k)
        s[name + ".bias"] = (co,)

    def norm(name, c):
        s[name + ".weight"] = s[name + ".bias"] = (c,)

    def res(p, ci, co):
        norm(p + "norm1", ci), conv(p + "conv1.conv", co, ci, 3), norm(p + "norm2", co), conv(p + "conv2.conv", co, co, 3)
        if ci != co:
            conv(p + "conv_shortcut.conv", co, ci, 1)
    zc = cfg["latent_channels"]
    blocks, c_mid = hyvae10_encoder_layout(cfg)
    conv("encoder.conv_in.conv", cfg["block_out_channels"][0], in_channels, 3)
    for i, (rs, down) in enumerate(blocks):
        for j, (ci, co) in enumerate(rs):
            res(f"encoder.down_blocks.{i}.resnets.{j}.", ci, co)
        if down is not None:
            conv(f"encoder.down_blocks.{i}.downsamplers.0.conv.conv", rs[-1][1], rs[-1][1], 3)
    res("encoder.mid_block.resnets.0.", c_mid, c_mid), res("encoder.mid_block.resnets.1.", c_mid, c_mid)
    a = "encoder.mid_block.attentions.0."
    norm(a + "group_norm", c_mid)
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        s[a + n + ".weight"], s[a + n + ".bias"] = (c_mid, c_mid), (c_mid,)
    norm("encoder.conv_norm_out", c_mid)
    conv("encoder.conv_out.conv", 2 * zc, c_mid, 3)
    conv("quant_conv", 2 * zc, 2 * zc, 1)
    return s


def make_hyvae10_state_dict(cfg, seed=0, device="cpu", dtype=torch.float32, encoder=False):
    out = {}
    shapes = dict(hyvae10_param_shapes(cfg))
    if encoder:
        shapes.update(hyvae10_encoder_param_shapes(cfg))
    for n, s in shapes.items():
        if "norm" in n and n.endswith(".weight"):
            out[n] = (1.0 + _normal(s, 0.1, seed, n, device)).to(dtype)
        else:
            out[n] = make_vae_tensor(n, s, seed, device).to(dtype)
    return out


# ---------------------------------------------------------------- umT5 text encoder (models/wan/modules/t5.py:459-472 umt5_xxl, encoder only)
T5_CONFIGS = {
    "umt5_xxl": dict(vocab_size=256384, dim=4096, dim_attn=4096, dim_ffn=10240, num_heads=64, num_layers=24, num_buckets=32),
    # reduced configs for parity tests (head dim stays 64 as in umT5-XXL)
    "t5_small": dict(vocab_size=1000, dim=256, dim_attn=256, dim_ffn=512, num_heads=4, num_layers=2, num_buckets=32),
    "t5_1layer_xxl": dict(vocab_size=2048, dim=4096, dim_attn=4096, dim_ffn=10240, num_heads=64, num_layers=1, num_buckets=32),
    # classic T5 v1.1 layout (shared relative position embedding): google/byt5-small, the glyph encoder of Hunyuan Video 1.5
    # (d_model 1472, 6 heads x 64, d_ff 3584, 12 layers; vocab 384 + the glyph colour / font tokens)
    "byt5_small": dict(vocab_size=1510, dim=1472, dim_attn=384, dim_ffn=3584, num_heads=6, num_layers=12, num_buckets=32, shared_pos=True),
    "byt5_tiny": dict(vocab_size=400, dim=192, dim_attn=128, dim_ffn=320, num_heads=2, num_layers=3, num_buckets=32, shared_pos=True),
    "byt5_2layer": dict(vocab_size=1510, dim=1472, dim_attn=384, dim_ffn=3584, num_heads=6, num_layers=2, num_buckets=32, shared_pos=True),
}


def t5_param_shapes(cfg):
    """Names and shapes of T5Encoder(shared_pos=False) (t5.py:268-281, 165-182, 75-90, 133-142)."""
    d, da, df, h, nb = cfg["dim"], cfg["dim_attn"], cfg["dim_ffn"], cfg["num_heads"], cfg["num_buckets"]
    shapes = {"token_embedding.weight": (cfg["vocab_size"], d), "norm.weight": (d,)}
    if cfg.get("shared_pos"):
        shapes["pos_embedding.embedding.weight"] = (nb, h)
    for i in range(cfg["num_layers"]):
        b = f"blocks.{i}."
        shapes.update({b + "norm1.weight": (d,), b + "attn.q.weight": (da, d), b + "attn.k.weight": (da, d), b + "attn.v.weight": (da, d),
                       b + "attn.o.weight": (d, da), b + "norm2.weight": (d,), b + "ffn.gate.0.weight": (df, d), b + "ffn.fc1.weight": (df, d),
                       b + "ffn.fc2.weight": (d, df)})
        if not cfg.get("shared_pos"):
            shapes[b + "pos_embedding.embedding.weight"] = (nb, h)
    return shapes


def make_t5_tensor(name, shape, cfg, seed=0, device="cpu"):
    """The reference's init_weights (t5.py:29-47) tempered so that a bf16-vs-fp32 comparison sees a realistic softmax: with the stock
    q std (dim * dim_attn)^-0.5 every logit is ~0 and the attention is uniform whatever the kernel does."""
    d, df = cfg["dim"], cfg["dim_ffn"]
    if name.endswith(("norm1.weight", "norm2.weight", "norm.weight")):
        return _normal(shape, 0.1, seed, name, device, mean=1.0)
    if name == "token_embedding.weight":
        return _normal(shape, 1.0, seed, name, device)
    if name.endswith("pos_embedding.embedding.weight"):
        return _normal(shape, 0.5, seed, name, device)
    if name.endswith("attn.q.weight"):
        return _normal(shape, 0.35 * d ** -0.5, seed, name, device)
    if name.endswith(("attn.k.weight", "attn.v.weight", "ffn.gate.0.weight", "ffn.fc1.weight")):
        return _normal(shape, d ** -0.5, seed, name, device)
    if name.endswith("attn.o.weight"):
        return _normal(shape, cfg["dim_attn"] ** -0.5, seed, name, device)
    if name.endswith("ffn.fc2.weight"):
        return _normal(shape, df ** -0.5, seed, name, device)
    raise KeyError(name)


def make_t5_state_dict(cfg, seed=0, device="cpu", dtype=torch.float32):
    return {k: make_t5_tensor(k, s, cfg, seed, device).to(dtype) for k, s in t5_param_shapes(cfg).items()}


def make_t5_inputs(cfg, length, n_valid, seed=0):
    """ids [length] (padding id 0 after n_valid tokens, as the reference tokenizer pads, tokenizers.py), mask [length]."""
    g = torch.Generator().manual_seed(1000 + seed)
    ids = torch.randint(1, cfg["vocab_size"], (length,), generator=g)
    ids[n_valid:] = 0
    mask = (torch.arange(length) < n_valid).long()
    return ids, mask


def t5_to_hf_t5stack_names(sd, num_layers):
    """Reference T5Encoder names -> transformers T5Stack names (the inverse of wan2gp_b200/wan/t5.py::hf_to_wan_names for the classic,
    shared-position layout: the bias lives in block 0)."""
    out = {"embed_tokens.weight": sd["token_embedding.weight"], "final_layer_norm.weight": sd["norm.weight"],
           "block.0.layer.0.SelfAttention.relative_attention_bias.weight": sd["pos_embedding.embedding.weight"]}
    for i in range(num_layers):
        b, h = f"blocks.{i}.", f"block.{i}.layer."
        out[h + "0.layer_norm.weight"], out[h + "1.layer_norm.weight"] = sd[b + "norm1.weight"], sd[b + "norm2.weight"]
        for n in "qkvo":
            out[h + f"0.SelfAttention.{n}.weight"] = sd[b + f"attn.{n}.weight"]
        out[h + "1.DenseReluDense.wi_0.weight"], out[h + "1.DenseReluDense.wi_1.weight"] = sd[b + "ffn.gate.0.weight"], sd[b + "ffn.fc1.weight"]
        out[h + "1.DenseReluDense.wo.weight"] = sd[b + "ffn.fc2.weight"]
    return out


# --------------------------------------------------------------------------- decoder-only LLM text towers (Hunyuan text encoders)

LLM_CONFIGS = {
    # Qwen2.5-VL-7B-Instruct language model (config.json of the checkpoint the reference downloads: hunyuan_handler.py:47-53)
    "qwen25_vl_7b": dict(vocab_size=152064, hidden_size=3584, intermediate_size=18944, num_layers=28, num_heads=28, num_kv_heads=4,
                         rms_eps=1e-6, rope_theta=1e6, qkv_bias=True),
    # llava-llama-3-8b language tower (Llama-3-8B + the llava tokens; hunyuan_handler.py:55-60)
    "llama3_8b": dict(vocab_size=128320, hidden_size=4096, intermediate_size=14336, num_layers=32, num_heads=32, num_kv_heads=8,
                      rms_eps=1e-5, rope_theta=5e5, qkv_bias=False),
    # reduced configs for parity tests (head dim stays 128)
    "qwen_tiny": dict(vocab_size=300, hidden_size=256, intermediate_size=512, num_layers=4, num_heads=2, num_kv_heads=1, rms_eps=1e-6,
                      rope_theta=1e6, qkv_bias=True),
    "llama_tiny": dict(vocab_size=300, hidden_size=512, intermediate_size=768, num_layers=3, num_heads=4, num_kv_heads=2, rms_eps=1e-5,
                       rope_theta=5e5, qkv_bias=False),
    "qwen_2layer_7b": dict(vocab_size=1024, hidden_size=3584, intermediate_size=18944, num_layers=2, num_heads=28, num_kv_heads=4,
                           rms_eps=1e-6, rope_theta=1e6, qkv_bias=True),
}


def llm_param_shapes(cfg):
    """transformers names (without the `model.` / `model.language_model.` prefix) and shapes of Qwen2_5_VLTextModel / LlamaModel."""
    D, Fi, H, Hk = cfg["hidden_size"], cfg["intermediate_size"], cfg["num_heads"], cfg["num_kv_heads"]
    s = {"embed_tokens.weight": (cfg["vocab_size"], D), "norm.weight": (D,)}
    for i in range(cfg["num_layers"]):
        p = f"layers.{i}."
        s.update({p + "input_layernorm.weight": (D,), p + "post_attention_layernorm.weight": (D,),
                  p + "self_attn.q_proj.weight": (H * 128, D), p + "self_attn.k_proj.weight": (Hk * 128, D),
                  p + "self_attn.v_proj.weight": (Hk * 128, D), p + "self_attn.o_proj.weight": (D, H * 128),
                  p + "mlp.gate_proj.weight": (Fi, D), p + "mlp.up_proj.weight": (Fi, D), p + "mlp.down_proj.weight": (D, Fi)})
        if cfg["qkv_bias"]:
            s.update({p + "self_attn.q_proj.bias": (H * 128,), p + "self_attn.k_proj.bias": (Hk * 128,), p + "self_attn.v_proj.bias": (Hk * 128,)})
    return s


def make_llm_tensor(name, shape, cfg, seed=0, device="cpu"):
    """Unit-variance activations everywhere (q . k / sqrt(128) ~ N(0, 1): a real softmax, not a uniform one)."""
    if name.endswith("layernorm.weight") or name == "norm.weight":
        return _normal(shape, 0.1, seed, name, device, mean=1.0)
    if name == "embed_tokens.weight":
        return _normal(shape, 1.0, seed, name, device)
    if name.endswith(".bias"):
        return _normal(shape, 0.1, seed, name, device)
    return _normal(shape, shape[1] ** -0.5, seed, name, device)


def make_llm_state_dict(cfg, seed=0, device="cpu", dtype=torch.float32):
    return {k: make_llm_tensor(k, s, cfg, seed, device).to(dtype) for k, s in llm_param_shapes(cfg).items()}


def make_llm_inputs(cfg, length, n_valid, seed=0):
    """ids [length] (right padded with id 0, as the reference tokenizers pad: padding_side="right"), mask [length]."""
    g = torch.Generator().manual_seed(2000 + seed)
    ids = torch.randint(1, cfg["vocab_size"], (length,), generator=g)
    ids[n_valid:] = 0
    return ids, (torch.arange(length) < n_valid).long()

