"""CPU dry-run of the drop-in boundary levels 1-2 for the Hunyuan family: the plugin's second model handler resolves
to a `family_handler` with the static methods WanGP's registry calls (hunyuan_handler.py:8-357), `load_model` returns
`(pipeline_obj, pipe_dict)` (:239-278) and `pipeline_obj.generate(**kwargs)` honours the keyword / return / callback / interrupt contract of
`HunyuanVideoSampler.generate` (hunyuan.py:728-1085) and the pipeline it drives (pipeline_hunyuan_video.py:1100-1830).  The C ABI is stubbed
(no arithmetic); the arithmetic behind generate() is covered by tests/test_hy_plugin_gpu.py."""
import importlib
import json
import os
import types

import pytest
import torch

from tests.test_host_dryrun_cpu import stub_abi  # noqa: F401  (fixture)
from wan2gp_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# reduced architectures whose latent channels match the reduced VAEs (synth.HYVAE_CONFIGS / HYVAE10_CONFIGS)
HY15_CFG = dict(synth.HY_CONFIGS["hy_tiny"], in_channels=17, out_channels=8)          # 8 latent + 8 condition + 1 mask channels
HY10_CFG = dict(synth.HY_CONFIGS["hy10_tiny"], in_channels=8, out_channels=8)
HY15_I2V_CFG = dict(HY15_CFG, vision_states_dim=64)                                      # + the vision_in projection of hunyuan_1_5_i2v
VAE15_KW = dict(latent_channels=8, block_out_channels=[32, 64, 64], layers_per_block=1, ffactor_spatial=4, ffactor_temporal=2, sample_size=16,
                sample_tsize=8, scaling_factor=1.03, shift_factor=0.1)


def load_hy_handler():
    info = json.load(open(os.path.join(ROOT, "plugin", "plugin_info.json")))
    mods = [m for m in info["model_handlers"] if "hunyuan" in m]
    assert len(mods) == 1 and mods[0].startswith(".")
    return importlib.import_module("plugin" + mods[0])


class FakeLLM:
    """The reference text-encoder protocol (text_encoder/__init__.py::TextEncoder, text_encoder_1_5.py): text2tokens -> encode ->
    `.hidden_state` [B, n, dim], `.attention_mask` [B, n] with a valid prefix.  Deterministic in the prompt text."""
    max_length, dtype = 12, torch.float32

    def __init__(self, dim, pooled=False):
        self.dim, self.pooled, self.calls = dim, pooled, []

    def text2tokens(self, prompts, data_type="video", max_length=None, name=None):
        self.calls.append(("text2tokens", tuple(prompts), data_type))
        return {"prompts": list(prompts)}

    def encode(self, tokens, data_type="video", device=None, is_uncond=False, semantic_images=None):
        self.calls.append(("encode", is_uncond))
        hs, ms = [], []
        for p in tokens["prompts"]:
            g = torch.Generator().manual_seed(len(p) + 17 * self.dim)
            if self.pooled:
                hs.append(torch.randn(self.dim, generator=g))
                continue
            n = 3 + len(p) % 5
            h = torch.randn(self.max_length, self.dim, generator=g)
            m = torch.zeros(self.max_length, dtype=torch.long)
            m[:n] = 1
            hs.append(h), ms.append(m)
        return types.SimpleNamespace(hidden_state=torch.stack(hs), attention_mask=None if self.pooled else torch.stack(ms))


class FakeVision:
    """WanGP's SigLIP `VisionEncoder` surface (vision_encoder/__init__.py:185-209): encode_images(uint8 HWC array) -> .last_hidden_state."""

    def __init__(self, dim, n_tokens=40):
        self.dim, self.n_tokens, self.seen = dim, n_tokens, []

    def encode_images(self, images):
        import numpy as np
        assert isinstance(images, np.ndarray) and images.dtype == np.uint8 and images.ndim == 3 and images.shape[2] == 3
        self.seen.append(images)
        g = torch.Generator().manual_seed(int(images.astype(np.int64).sum()) % 100003)
        return types.SimpleNamespace(last_hidden_state=torch.randn(1, self.n_tokens, self.dim, generator=g))


def hy_kwargs(**over):
    """The keyword set wgp.py passes to a Hunyuan pipeline object (the common generate(**kwargs) call, wgp.py:7762-7880): the named
    arguments of hunyuan.py:728-757 plus the ones that land in **kwargs."""
    kw = dict(input_prompt="a red fox", input_ref_images=None, audio_guide=None, input_frames=None, input_masks=None, input_video=None, fps=24,
              height=32, width=48, frame_num=5, seed=3, n_prompt="", sampling_steps=3, guide_scale=6.0, shift=9.0, embedded_guidance_scale=6.0,
              batch_size=1, image_start=None, enable_RIFLEx=False, VAE_tile_size=None, joint_pass=False, cfg_star_switch=0,
              fit_into_canvas=True, conditioning_latents_size=0, callback=None, alt_prompt="", sample_solver="", model_type="b200_hunyuan_1_5_t2v",
              loras_slists=None, image_mode=0, video_prompt_type="", window_no=1, offloadobj=None, set_header_text=lambda *_: None,
              NAG_scale=1, model_mode=None, guide2_scale=3.0, switch_threshold=0, guide_phases=1)
    kw.update(over)
    return kw


def make_pipeline(arch, device="cpu", monkeypatch=None, **extra):
    h = load_hy_handler()
    v15, i2v = arch.startswith("b200_hunyuan_1_5"), arch.endswith("i2v")
    cfg = (HY15_I2V_CFG if i2v else HY15_CFG) if v15 else HY10_CFG
    monkeypatch.setitem(synth.HY_CONFIGS, h.ARCHS[arch][0], cfg)          # reduced architecture behind the production name
    sd = synth.make_hy_state_dict(cfg, 0)
    if v15:
        vcfg = synth.HYVAE_CONFIGS["hyvae_tiny"]
        vsd = {"decoder." + k: v for k, v in synth.make_hyvae_state_dict(vcfg, 0).items()}
        vae_cfg = VAE15_KW
        te = dict(text_encoder=FakeLLM(cfg["text_states_dim"]))
        if i2v:
            vsd.update({"encoder." + k: v for k, v in synth.make_hyvae_state_dict(vcfg, 0, encoder=True).items()})
            te["vision_encoder"] = FakeVision(cfg["vision_states_dim"])
    else:
        vae_cfg = dict(synth.HYVAE10_CONFIGS["hyvae10_tiny"], sample_size=32, sample_tsize=16, scaling_factor=0.476986)
        vsd = synth.make_hyvae10_state_dict(synth.HYVAE10_CONFIGS["hyvae10_tiny"], 0)
        te = dict(text_encoder=FakeLLM(cfg["text_states_dim"]), text_encoder_2=FakeLLM(cfg["text_states_dim_2"], pooled=True))
    model_def = json.load(open(os.path.join(ROOT, "plugin", "defaults", arch + ".json")))["model"]
    pipe_obj, pipe = h.family_handler.load_model(["dit.safetensors"], arch, arch, model_def, dtype=torch.bfloat16, VAE_dtype=torch.float32,
                                                 text_encoder_filename=None, profile=1, state_dict=sd, vae_state_dict=vsd, vae_cfg=vae_cfg,
                                                 device=device, **te, **extra)
    return pipe_obj, pipe, cfg, sd, vsd


def test_hunyuan_family_handler_contract():
    fh = load_hy_handler().family_handler
    for name in ("query_supported_types", "query_family_maps", "query_model_family", "query_family_infos", "query_model_def",
                 "query_model_files", "load_model", "fix_settings", "update_default_settings", "validate_generative_settings",
                 "set_cache_parameters", "get_lora_dir", "register_lora_cli_args", "get_rgb_factors"):
        assert callable(getattr(fh, name)), name                    # hunyuan_handler.py:8-357 / wgp.py:2717-2735
    types_ = fh.query_supported_types()
    assert set(types_) == {"b200_hunyuan_1_5_t2v", "b200_hunyuan_1_5_i2v", "b200_hunyuan"}
    for t in types_:
        d = json.load(open(os.path.join(ROOT, "plugin", "defaults", t + ".json")))
        assert d["model"]["architecture"] == t and all("quanto" not in u for u in d["model"]["URLs"])
        md = fh.query_model_def(t, d["model"])
        assert md["fps"] == 24 and md["frames_steps"] == 4 and md["flow_shift"]
        assert ("embedded_guidance" in md) == (t == "b200_hunyuan")   # hunyuan_handler.py:82-85
        assert fh.query_model_files([], t, d["model"])
    assert fh.query_model_family() in fh.query_family_infos()
    ui = {}
    fh.update_default_settings("b200_hunyuan", {}, ui)
    assert ui == {"embedded_guidance_scale": 6.0, "guidance_scale": 7.0}
    assert fh.validate_generative_settings("b200_hunyuan", {}, {"activated_loras": ["x"]}) is not None
    assert fh.validate_generative_settings("b200_hunyuan", {}, {"image_prompt_type": "S"}) is not None
    assert fh.validate_generative_settings("b200_hunyuan_1_5_i2v", {}, {"image_prompt_type": "S"}) is None
    assert fh.query_model_def("b200_hunyuan_1_5_i2v", {})["i2v_class"] and "siglip_vision_model" in fh.query_model_files([], "b200_hunyuan_1_5_i2v")[1]["sourceFolderList"]
    assert fh.validate_generative_settings("b200_hunyuan_1_5_t2v", {}, {}) is None
    with pytest.raises(NotImplementedError):
        fh.set_cache_parameters("tea", "b200_hunyuan", {}, {}, None)
    # both handlers of the plugin register disjoint model types
    wan = importlib.import_module("plugin.models.b200_wan_handler").family_handler
    assert not set(types_) & set(wan.query_supported_types())


def test_generate_hunyuan_1_5_contract(stub_abi, monkeypatch):  # noqa: F811
    pipe_obj, pipe, cfg, _, _ = make_pipeline("b200_hunyuan_1_5_t2v", monkeypatch=monkeypatch)
    assert {"transformer", "vae"} <= set(pipe) and all(isinstance(m, torch.nn.Module) for m in pipe.values())
    assert pipe_obj.model is pipe["transformer"] and pipe_obj.vae is pipe["vae"] and pipe_obj._interrupt is False
    assert pipe_obj.hunyuan_1_5 and pipe_obj.enable_cfg and pipe_obj.pipeline._interrupt is False
    events = []

    def callback(step=-1, latents=None, force=True, read_state=False, override_num_inference_steps=-1, pass_no=-1, denoising_extra=""):
        events.append((step, None if latents is None else tuple(latents.shape), force, read_state))
    out = pipe_obj.generate(**hy_kwargs(callback=callback))
    assert out.dtype == torch.float32 and out.device.type == "cpu" and tuple(out.shape) == (3, 5, 32, 48)
    # callback protocol: (-1, None, True) once, per-block polls (-1, None, False, True) from the transformer, (i, latents, False) per step
    assert events[0][:3] == (-1, None, True)
    steps = [e for e in events if e[0] >= 0]
    assert [e[0] for e in steps] == [0, 1, 2] and all(e[1] == (8, 3, 8, 12) and e[2] is False for e in steps)
    n_blocks = cfg["mm_double_blocks_depth"]
    assert sum(1 for e in events if e[0] == -1 and e[3] is True) == 3 * 2 * n_blocks           # 3 steps x CFG pair x blocks
    # CFG pair: negative prompt encoded with is_uncond=True (encode_prompt_1_5 :394-400); no guidance -> one forward, no negative encode
    te = pipe_obj.text_encoder
    assert ("encode", True) in te.calls
    te.calls.clear()
    n0 = stub_abi.count("b200_cfg_euler_step")
    assert pipe_obj.generate(**hy_kwargs(guide_scale=1.0, joint_pass=True)).shape == (3, 5, 32, 48)
    assert ("encode", True) not in te.calls and stub_abi.count("b200_cfg_euler_step") - n0 == 3
    # joint pass (one forward of batch 2) and CFG-Zero*, seed list, the reference's tile settings
    out = pipe_obj.generate(**hy_kwargs(joint_pass=True, cfg_star_switch=1, seed=[5], frame_num=9, height=48, width=64,
                                        VAE_tile_size={"tile_sample_min_size": 16, "tile_sample_min_tsize": 8}))
    assert tuple(out.shape) == (3, 9, 48, 64) and pipe_obj.vae.use_spatial_tiling and pipe_obj.vae.tile_sample_min_size == 16
    with pytest.raises(ValueError):
        pipe_obj.generate(**hy_kwargs(frame_num=6))
    with pytest.raises(TypeError):
        pipe_obj.generate(**hy_kwargs(input_prompt=["a", "b"]))
    for bad in (dict(image_start=torch.zeros(3, 32, 48)), dict(input_ref_images=[object()]), dict(input_video=torch.zeros(3, 5, 32, 48)),
                dict(audio_guide="a.wav")):
        with pytest.raises(NotImplementedError):
            pipe_obj.generate(**hy_kwargs(**bad))


def test_generate_hunyuan_1_0_contract_and_interrupt(stub_abi, monkeypatch):  # noqa: F811
    pipe_obj, pipe, cfg, _, _ = make_pipeline("b200_hunyuan", monkeypatch=monkeypatch)
    assert not pipe_obj.hunyuan_1_5 and not pipe_obj.enable_cfg and pipe_obj.model.guidance_embed
    n0 = stub_abi.count("b200_cfg_euler_step")
    out = pipe_obj.generate(**hy_kwargs(model_type="b200_hunyuan", guide_scale=7.0, height=32, width=48, frame_num=9, shift=7.0))
    assert tuple(out.shape) == (3, 9, 32, 48) and stub_abi.count("b200_cfg_euler_step") - n0 == 3
    assert ("encode", True) not in pipe_obj.text_encoder.calls       # guidance-distilled: guide_scale is forced to 1, no negative branch
    n = {"polls": 0}

    def callback(step=-1, latents=None, force=True, read_state=False, **kw):
        n["polls"] += 1
        if n["polls"] == 4:
            pipe_obj._interrupt = True                                # the UI thread's abort
    assert pipe_obj.generate(**hy_kwargs(model_type="b200_hunyuan", callback=callback)) is None
    pipe_obj._interrupt = False
    assert pipe_obj.generate(**hy_kwargs(model_type="b200_hunyuan")) is not None


def test_generate_hunyuan_1_5_i2v_contract(stub_abi, monkeypatch):  # noqa: F811
    """hunyuan_1_5_i2v: the start frame is VAE-encoded into frame 0 of the concat condition (mask 1 there), the image encoder sees the uint8
    frame (convert_tensor_to_image), its tokens go into every forward; geometry follows the image (hunyuan.py:879-893, pipeline :1513-1531)."""
    import wan2gp_b200.pipeline as pl
    pipe_obj, pipe, cfg, _, _ = make_pipeline("b200_hunyuan_1_5_i2v", monkeypatch=monkeypatch)
    assert pipe_obj.i2v_mode and pipe_obj.default_negative_prompt.startswith("deformation") and "vision_encoder" not in pipe   # not an nn.Module here
    seen = []
    real_step = pl.HunyuanDenoiser.step

    def spy(self, latents, cond_latents, i, *a, **k):
        seen.append((cond_latents.clone(), k.get("vision_states")))
        return real_step(self, latents, cond_latents, i, *a, **k)
    monkeypatch.setattr(pl.HunyuanDenoiser, "step", spy)
    img = torch.rand(3, 32, 48, generator=torch.Generator().manual_seed(1)) * 2 - 1
    out = pipe_obj.generate(**hy_kwargs(model_type="b200_hunyuan_1_5_i2v", image_start=img, height=64, width=64, shift=7.0))
    assert tuple(out.shape) == (3, 5, 32, 48)                                       # the image's size wins over height / width
    cond, vis = seen[0]
    assert tuple(cond.shape) == (1, 9, 3, 8, 12) and float(cond[:, 8, 0].min()) == 1.0 and float(cond[:, 8, 1:].abs().max()) == 0.0
    assert float(cond[:, :8, 1:].abs().max()) == 0.0 and tuple(vis.shape) == (1, 40, 64) and vis.dtype == torch.bfloat16
    ve = pipe_obj.vision_encoder
    want = img.clone().add_(1.).mul_(127.5).permute(1, 2, 0).to(torch.uint8).numpy()
    assert len(ve.seen) == 1 and (ve.seen[0] == want).all()
    assert {"b200_hy_downsample_cl", "b200_group_mean_cl"} <= set(stub_abi)          # the VAE encoder ran
    # the first frame of a video works the same way; without either the call is an error; t2v objects still reject images
    out = pipe_obj.generate(**hy_kwargs(model_type="b200_hunyuan_1_5_i2v", input_video=torch.rand(3, 2, 32, 48) * 2 - 1, joint_pass=True))
    assert tuple(out.shape) == (3, 5, 32, 48)
    with pytest.raises(ValueError):
        pipe_obj.generate(**hy_kwargs(model_type="b200_hunyuan_1_5_i2v"))
    with pytest.raises(NotImplementedError):
        load_hy_handler()  # noqa: B018
        from wan2gp_b200.hyvideo.hunyuan import HunyuanVideoSampler
        HunyuanVideoSampler(pipe_obj.model, pipe_obj.vae, hunyuan_1_5=False, i2v=True, device="cpu")

