"""Family handler of the drop-in boundary, level 1: registers the B200-native Wan path with WanGP's model
registry through the plugin API (`plugin_info.json: {"type": "model", "model_handlers": [...]}`, docs/PLUGINS.md:44-58;
shared/utils/plugins.py:266-271, 671-698; wgp.py:2649-2657, 2717-2735) -- the same `family_handler` contract as the built-in
`models/wan/wan_handler.py` (:72-1500), restricted to the model types of the hot path:

    b200_t2v_1.3B, b200_t2v_2_2 (two experts), b200_i2v_2_2 (two experts, start-image / prefix-video conditioning)

`load_model` returns `(pipeline_obj, pipe_dict)` like wan_handler.py:1117-1158: `pipeline_obj` is `wan2gp_b200.wan.WanAny2V`
(level 2: `generate(**kwargs)`, `_interrupt`, `.model`, `.model2`, `.vae`), `pipe_dict` maps the mmgp component names to
`nn.Module`s (`transformer`, `transformer2`, `vae`, `text_encoder`).  The text encoder is the umT5 of `wan2gp_b200/wan/t5.py` (WanGP's own `T5EncoderModel` for quantised encoder files), it
sits in front of the hot path; the transformer(s) and the VAE are the sm_100a implementations.  There is no CPU / eager fallback:
loading on a machine without the CUDA library or an sm_100 device raises."""
import os

import torch

ARCHS = {   # plugin architecture name -> (wan2gp_b200.synth.WAN_CONFIGS key, two experts?)
    "b200_t2v_1.3B": ("t2v_1.3B", False),
    "b200_t2v_2_2": ("t2v_2_2", True),
    "b200_i2v_2_2": ("i2v_2_2", True),
}


def _read_state_dict(path):
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file
        return load_file(path, device="cpu")
    return torch.load(path, map_location="cpu")


class family_handler:
    @staticmethod
    def query_supported_types():
        return list(ARCHS)

    @staticmethod
    def query_family_maps():
        return {}, {}                                        # no equivalence / compatibility classes with the built-in types

    @staticmethod
    def query_model_family():
        return "wan_b200"

    @staticmethod
    def query_family_infos():
        return {"wan_b200": (2, "Wan (B200 native)")}

    @staticmethod
    def register_lora_cli_args(parser, lora_root):
        return None                                          # LoRAs are applied by mmgp hooks on nn.Linear: outside this path

    @staticmethod
    def get_lora_dir(base_model_type, args, lora_root):
        return os.path.join(lora_root, "wan_b200")

    @staticmethod
    def set_cache_parameters(cache_type, base_model_type, model_def, inputs, skip_steps_cache):
        raise NotImplementedError("TeaCache / MagCache step skipping is outside the B200 hot path")

    @staticmethod
    def query_model_def(base_model_type, model_def):
        """Capabilities the UI reads (subset of wan_handler.py:217-1008 that applies to plain t2v / i2v_2_2)."""
        i2v = base_model_type == "b200_i2v_2_2"
        folder = "umt5-xxl"
        extra = {
            "riflex": True, "i2v_class": i2v, "t2v_class": not i2v, "i2v_2_2": i2v, "color_correction": False,
            "text_encoder_folder": model_def.get("text_encoder_folder", folder),
            "text_encoder_URLs": model_def.get("text_encoder_URLs", [
                f"https://huggingface.co/DeepBeepMeep/Wan2.1/resolve/main/{folder}/models_t5_umt5-xxl-enc-bf16.safetensors"]),
            "vae_block_size": 16, "fps": 16, "frames_minimum": 5, "frames_steps": 4, "sliding_window": False,
            "multiple_submodels": "URLs2" in model_def, "guidance_max_phases": 3, "perturbation": False, "flow_shift": True,
            "cfg_zero": True, "cfg_star": True, "adaptive_projected_guidance": False, "tea_cache": False, "mag_cache": False,
            "no_steps_skipping": True, "self_refiner": False, "group": "wan_b200", "profiles_dir": [],
            "sample_solvers": [("unipc", "unipc"), ("euler", "euler"), ("dpm++", "dpm++"), ("flowmatch causvid", "causvid"), ("lcm + ltx", "lcm")],
        }
        if i2v:
            extra["motion_amplitude"] = True
            extra["image_prompt_types_allowed"] = "SV"       # start image / continue video
        return extra

    @staticmethod
    def get_rgb_factors(base_model_type):
        try:                                                 # latent preview colours: WanGP's own table for the Wan2.1 VAE
            from shared.RGB_factors import get_rgb_factors
            return get_rgb_factors("wan", "t2v")
        except ImportError:
            return None, None

    @staticmethod
    def query_model_files(computeList, base_model_type, model_def=None):
        """Tokenizer + VAE downloads, as wan_handler.py:1017-1048 for the non-5B models (no CLIP: t2v / i2v_2_2 do not use it)."""
        return [{"repoId": "DeepBeepMeep/Wan2.1", "sourceFolderList": ["umt5-xxl", ""],
                 "fileList": [["special_tokens_map.json", "spiece.model", "tokenizer.json", "tokenizer_config.json"], ["Wan2.1_VAE.safetensors"]]}]

    @staticmethod
    def load_model(model_filename, model_type, base_model_type, model_def, quantizeTransformer=False, text_encoder_quantization=None,
                   dtype=torch.bfloat16, VAE_dtype=torch.float32, mixed_precision_transformer=False, save_quantized=False,
                   submodel_no_list=None, text_encoder_filename=None, VAE_upsampling=None, text_encoder=None, vae_state_dict=None,
                   vae_cfg=None, state_dicts=None, device="cuda", **kwargs):
        """-> (pipeline_obj, pipe_dict).  `model_filename`: list of checkpoint paths (high-noise expert first, wan_handler.py /
        any2video.py:170-232).  Extra keyword-only hooks for tests / embedders: `state_dicts` (already loaded weights instead of
        files), `vae_state_dict` / `vae_cfg` (reduced VAE in tests), `text_encoder` (callable(prompts, device) -> list of [len, 4096] tensors)."""
        from wan2gp_b200 import synth
        from wan2gp_b200.wan import WanAny2V, WanModel, WanVAE
        if base_model_type not in ARCHS:
            raise ValueError(f"b200_wan_handler: unsupported model type {base_model_type!r}")
        if quantizeTransformer or save_quantized:
            raise NotImplementedError("quantised transformer weights are outside the B200 bf16 hot path")
        cfg_name, two = ARCHS[base_model_type]
        cfg = synth.WAN_CONFIGS[cfg_name]
        files = [model_filename] if isinstance(model_filename, str) else list(model_filename or [])
        sds = list(state_dicts) if state_dicts is not None else [_read_state_dict(f) for f in files[:2 if two else 1]]
        if len(sds) < (2 if two else 1):
            raise ValueError(f"{base_model_type}: expected {2 if two else 1} transformer checkpoint(s), got {len(sds)}")
        models = []
        for sd in sds[:2 if two else 1]:
            m = WanModel(**cfg, device=device)
            m.load_state_dict(sd)
            models.append(m)
        if vae_state_dict is None:
            vae_path = None
            try:
                from shared.utils import files_locator as fl            # WanGP's checkpoint locator
                vae_path = fl.locate_file("Wan2.1_VAE.safetensors")
            except ImportError:
                pass
            if vae_path is None:
                raise FileNotFoundError("Wan2.1_VAE.safetensors not found (query_model_files lists it for download)")
            vae_state_dict = _read_state_dict(vae_path)
        vae = WanVAE(device=device, state_dict=vae_state_dict, cfg=vae_cfg)
        te_module = None
        if text_encoder is None:
            # the umT5 encoder on the B200 kernels (wan2gp_b200/wan/t5.py), constructed the way any2video.py:119-126 constructs WanGP's own
            # T5EncoderModel -- same checkpoint (Wan or Hugging Face umT5 names), same tokenizer folder.  Quantised encoder files
            # (text_encoder_quantization) keep WanGP's encoder: they are not bf16 weights.
            tok = os.path.dirname(text_encoder_filename)
            if text_encoder_quantization in (None, "", "bf16"):
                from wan2gp_b200.wan.t5 import T5EncoderModel
                text_encoder = T5EncoderModel(text_len=cfg["text_len"], dtype=torch.bfloat16, device=device,
                                              checkpoint_path=text_encoder_filename, tokenizer_path=tok)
            else:
                from models.wan.modules.t5 import T5EncoderModel
                text_encoder = T5EncoderModel(text_len=cfg["text_len"], dtype=torch.bfloat16, device=torch.device("cpu"),
                                              checkpoint_path=text_encoder_filename, tokenizer_path=tok)
            te_module = text_encoder.model
        pipe_obj = WanAny2V(models[0], models[1] if two else None, vae, text_encoder, model_def=dict(model_def or {}, i2v_2_2=cfg["in_dim"] > 16),
                            base_model_type=base_model_type, device=device, dtype=dtype, VAE_dtype=VAE_dtype)
        pipe = {"transformer": models[0], "vae": vae.model}
        if two:
            pipe["transformer2"] = models[1]
        if te_module is not None:
            pipe["text_encoder"] = te_module
        return pipe_obj, pipe

    @staticmethod
    def fix_settings(base_model_type, settings_version, model_def, ui_defaults):
        if ui_defaults.get("sample_solver", "") == "":
            ui_defaults["sample_solver"] = "unipc"                           # wan_handler.py:1162-1163

    @staticmethod
    def update_default_settings(base_model_type, model_def, ui_defaults):
        ui_defaults.update({"sample_solver": "unipc"})
        if base_model_type == "b200_i2v_2_2":
            ui_defaults.update({"image_prompt_type": "S"})

    @staticmethod
    def validate_generative_settings(base_model_type, model_def, inputs):
        """Returns an error string for settings that would leave the hot path (WanGP shows it instead of queueing), else None."""
        if inputs.get("skip_steps_cache_type", "") not in ("", None):
            return "Step skipping (TeaCache / MagCache) is not available with the B200-native Wan path"
        if inputs.get("NAG_scale", 1) > 1:
            return "NAG is not available with the B200-native Wan path"
        if inputs.get("activated_loras"):
            return "LoRAs are not available with the B200-native Wan path"
        return None
