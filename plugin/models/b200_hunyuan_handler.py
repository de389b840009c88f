"""Family handler of the drop-in boundary, level 1 for the Hunyuan family: registers the B200-native Hunyuan
Video 1.5 / HunyuanVideo 1.0 text-to-video path with WanGP's model registry through the plugin API -- the same `family_handler`
contract as the built-in `models/hyvideo/hunyuan_handler.py` (:8-357), restricted to the model types of the hot path:

    b200_hunyuan_1_5_t2v  ('HYVideo-1_5': 54 double-stream blocks, 65 -> 32 channels, CFG pair, Hunyuan 1.5 VAE 16x / 4x)
    b200_hunyuan_1_5_i2v  (the same architecture; start image as latent-concat condition + image-encoder tokens, hunyuan.py:211-215)
    b200_hunyuan          ('HYVideo-T/2-cfgdistill': 20 double + 40 single blocks, embedded guidance, HunyuanVideo 1.0 VAE 8x / 4x)

`load_model` returns `(pipeline_obj, pipe_dict)` like hunyuan_handler.py:239-278: `pipeline_obj` is
`wan2gp_b200.hyvideo.hunyuan.HunyuanVideoSampler` (level 2: `generate(**kwargs)`, `_interrupt`, `.model`, `.vae`), `pipe_dict` maps the
mmgp component names to `nn.Module`s.  The text encoders (Qwen2.5-VL / llava-llama-3, CLIP-L, glyph-byT5) sit in front of the hot path
and are WanGP's own `TextEncoder` objects, built exactly as hunyuan.py:283-305, 372-446 builds them -- with the language model and the byT5
model inside them moved onto the B200 kernels (`wan2gp_b200.hyvideo.llm.LlamaLikeTextModel`, `.byt5.ByT5Encoder`) for bf16 encoder files;
the transformer and the VAE are the sm_100a implementations.  There is no CPU / eager fallback: loading without the CUDA library or an sm_100 device raises."""
import json
import os

import torch

ARCHS = {   # plugin architecture name -> (wan2gp_b200.synth.HY_CONFIGS key, Hunyuan 1.5?)
    "b200_hunyuan_1_5_t2v": ("HYVideo-1_5", True),
    "b200_hunyuan_1_5_i2v": ("HYVideo-1_5", True),
    "b200_hunyuan": ("HYVideo-T/2-cfgdistill", False),
}
I2V = {"b200_hunyuan_1_5_i2v"}          # latent-concat start image + SigLIP tokens through the model's vision_in projection


def _read_state_dict(path):
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file
        return load_file(path, device="cpu")
    return torch.load(path, map_location="cpu")


def _is_1_5(base_model_type):
    return ARCHS[base_model_type][1]


def build_transformer(cfg, v15, device="cuda", i2v=False):
    """HYVideoDiffusionTransformer with the constructor arguments hunyuan.py:227-253 / models.py:1280-1363 give the two families."""
    from wan2gp_b200.hyvideo import HYVideoDiffusionTransformer
    common = dict(i2v_condition_type="latent_concat" if i2v else None, patch_size=cfg["patch_size"], in_channels=cfg["in_channels"], out_channels=cfg["out_channels"],
                  hidden_size=cfg["hidden_size"], heads_num=cfg["heads_num"], mlp_width_ratio=cfg["mlp_width_ratio"],
                  mm_double_blocks_depth=cfg["mm_double_blocks_depth"], rope_dim_list=cfg["rope_dim_list"],
                  text_states_dim=cfg["text_states_dim"], device=device)
    if v15:
        return HYVideoDiffusionTransformer(mm_single_blocks_depth=0, text_pool_type=None, text_states_dim_2=None, glyph_byT5_v2=True,
                                           use_cond_type_embedding=True, pre_split_qkv=True, vision_projection="linear",
                                           vision_states_dim=cfg.get("vision_states_dim", 1152), **common)
    return HYVideoDiffusionTransformer(mm_single_blocks_depth=cfg["mm_single_blocks_depth"], text_states_dim_2=cfg["text_states_dim_2"],
                                       guidance_embed=bool(cfg.get("guidance_embed", False)), **common)


class family_handler:
    @staticmethod
    def query_supported_types():
        return list(ARCHS)

    @staticmethod
    def query_family_maps():
        return {}, {}

    @staticmethod
    def query_model_family():
        return "hunyuan_b200"

    @staticmethod
    def query_family_infos():
        return {"hunyuan_b200": (22, "Hunyuan Video (B200 native)")}

    @staticmethod
    def register_lora_cli_args(parser, lora_root):
        return None

    @staticmethod
    def get_lora_dir(base_model_type, args, lora_root):
        return os.path.join(lora_root, "hunyuan_b200")

    @staticmethod
    def set_cache_parameters(cache_type, base_model_type, model_def, inputs, skip_steps_cache):
        raise NotImplementedError("TeaCache / MagCache step skipping is outside the B200 hot path")

    @staticmethod
    def query_model_def(base_model_type, model_def):
        """Capabilities the UI reads (subset of hunyuan_handler.py:38-141 that applies to plain t2v)."""
        v15 = _is_1_5(base_model_type)
        folder = "Qwen2.5-VL-7B-Instruct" if v15 else "llava-llama-3-8b"
        urls = ([f"https://huggingface.co/DeepBeepMeep/Qwen_image/resolve/main/{folder}/Qwen2.5-VL-7B-Instruct_bf16.safetensors"] if v15 else
                [f"https://huggingface.co/DeepBeepMeep/HunyuanVideo/resolve/main/{folder}/llava-llama-3-8b-v1_1_vlm_fp16.safetensors"])
        extra = {"riflex": True, "text_encoder_folder": model_def.get("text_encoder_folder", folder),
                 "text_encoder_URLs": model_def.get("text_encoder_URLs", urls), "fps": 24, "frames_minimum": 5, "frames_steps": 4,
                 "sliding_window": False, "flow_shift": True, "cfg_star": v15, "tea_cache": False, "mag_cache": False,
                 "no_steps_skipping": True, "group": "hunyuan_b200", "profiles_dir": []}
        if base_model_type in I2V:
            extra.update({"i2v_class": True, "image_prompt_types_allowed": "S"})
        if v15:
            extra["guidance_max_phases"] = 1
        else:
            extra["embedded_guidance"] = True                                        # hunyuan_handler.py:82-83
        return extra

    @staticmethod
    def get_rgb_factors(base_model_type):
        try:
            from shared.RGB_factors import get_rgb_factors
            return get_rgb_factors("hunyuan", sub_family="hunyuan1.5" if _is_1_5(base_model_type) else "")
        except ImportError:
            return None, None

    @staticmethod
    def query_model_files(computeList, base_model_type, model_def=None):
        """The downloads hunyuan_handler.py:208-236 lists for the family, minus the files of the conditioning variants."""
        if _is_1_5(base_model_type):
            siglip = [["siglip_vision_model"], [["model.safetensors", "config.json", "preprocessor_config.json"]]] if base_model_type in I2V else [[], []]
            return [{"repoId": "DeepBeepMeep/Qwen_image", "sourceFolderList": ["Qwen2.5-VL-7B-Instruct"],
                     "fileList": [["merges.txt", "tokenizer_config.json", "config.json", "vocab.json", "video_preprocessor_config.json",
                                   "preprocessor_config.json", "chat_template.json"]]},
                    {"repoId": "DeepBeepMeep/HunyuanVideo1.5", "sourceFolderList": ["Glyph-SDXL-v2", "Glyph-SDXL-v2/byt5-small", ""] + siglip[0],
                     "fileList": [["color_idx.json", "multilingual_10-lang_idx.json"], ["config.json", "model.safetensors", "byt5_model.safetensors"],
                                  ["hunyuan_video_1_5_VAE_fp32.safetensors", "hunyuan_video_1_5_VAE.json"]] + siglip[1]}]
        return {"repoId": "DeepBeepMeep/HunyuanVideo", "sourceFolderList": ["llava-llama-3-8b", "clip_vit_large_patch14", ""],
                "fileList": [["config.json", "special_tokens_map.json", "tokenizer.json", "tokenizer_config.json", "preprocessor_config.json"],
                             ["config.json", "merges.txt", "model.safetensors", "preprocessor_config.json", "special_tokens_map.json",
                              "tokenizer.json", "tokenizer_config.json", "vocab.json"],
                             ["hunyuan_video_VAE_fp32.safetensors", "hunyuan_video_VAE_config.json"]]}

    @staticmethod
    def _reference_text_encoders(v15, model_def, text_encoder_filename, device, text_encoder_quantization=None):
        """The encoders in front of the path, built as hunyuan.py:283-305 (glyph-byT5) and :372-446 (LLM, CLIP-L) build them.
        Needs WanGP's own packages (`models.hyvideo`, `shared`): this runs inside WanGP.  WanGP's `TextEncoder` objects keep their
        tokenizer, prompt templates and crop logic; for bf16 encoder files the language model inside (`te.model`: transformers'
        Qwen2.5-VL / llava-llama-3) is replaced by `LlamaLikeTextModel` on the B200 kernels (same weights, same call surface), as the glyph
        byT5 model is replaced by `ByT5Encoder`.  Quantised encoder files keep WanGP's own modules; CLIP-L (HunyuanVideo 1.0) stays WanGP's."""
        from models.hyvideo.constants import PROMPT_TEMPLATE
        from shared.utils import files_locator as fl
        text_len = 512                                                               # hunyuan.py:179 (256 only for the avatar variant)
        video_tpl = "li-dit-encode-video-json" if v15 else "dit-llm-encode-video"
        image_tpl = "li-dit-encode-image-json" if v15 else "dit-llm-encode"
        max_length = text_len + PROMPT_TEMPLATE[video_tpl].get("crop_start", 0)
        folder = (model_def or {}).get("text_encoder_folder")
        if folder:
            tok_path = os.path.dirname(fl.locate_file(os.path.join(folder, "tokenizer_config.json")))
        else:
            tok_path = os.path.dirname(text_encoder_filename) if text_encoder_filename else None
        if v15:
            from models.hyvideo.text_encoder.text_encoder_1_5 import TextEncoder as TextEncoderCls
        else:
            from models.hyvideo.text_encoder import TextEncoder as TextEncoderCls
        te = TextEncoderCls(text_encoder_type="llm", max_length=max_length, text_encoder_precision="fp16", tokenizer_type="llm",
                            tokenizer_path=tok_path, i2v_mode=False, prompt_template=PROMPT_TEMPLATE[image_tpl],
                            prompt_template_video=PROMPT_TEMPLATE[video_tpl], hidden_state_skip_layer=2, apply_final_norm=False,
                            reproduce=True, device="cpu", image_embed_interleave=1, text_encoder_path=text_encoder_filename)
        if text_encoder_quantization in (None, "", "bf16"):
            from wan2gp_b200.hyvideo.llm import LLAMA3_8B, QWEN25_VL_7B, LlamaLikeTextModel
            preset = QWEN25_VL_7B if v15 else LLAMA3_8B
            te.model = LlamaLikeTextModel.from_state_dict(te.model.state_dict(), preset["num_heads"], preset["num_kv_heads"], preset["rms_eps"],
                                                          preset["rope_theta"], device=device)
        te2 = byt5_model = byt5_tok = fmt = None
        if v15:
            from models.hyvideo.text_encoder.byT5 import load_glyph_byT5_v2
            from models.hyvideo.text_encoder.byT5.format_prompt import MultilingualPromptFormat
            font = fl.locate_file("Glyph-SDXL-v2/color_idx.json")                   # (sic) hunyuan.py:289-290 names them crosswise
            color = fl.locate_file("Glyph-SDXL-v2/multilingual_10-lang_idx.json")
            byt5_tok, byt5_model = load_glyph_byT5_v2(dict(byT5_google_path=fl.locate_folder("Glyph-SDXL-v2/byt5-small"),
                                                           byT5_ckpt_path=fl.locate_file("Glyph-SDXL-v2/byt5-small/byt5_model.safetensors"),
                                                           multilingual_prompt_format_color_path=font, multilingual_prompt_format_font_path=color,
                                                           byt5_max_length=256), device=device)
            fmt = MultilingualPromptFormat(font_path=font, color_path=color)
            # the glyph encoder itself runs on the B200 kernels: same weights (the T5Stack WanGP just loaded), same call surface
            from wan2gp_b200.hyvideo.byt5 import ByT5Encoder
            byt5_model = ByT5Encoder.from_state_dict(byt5_model.state_dict(), device=device)
        else:
            from models.hyvideo.text_encoder import TextEncoder
            te2 = TextEncoder(text_encoder_type="clipL", max_length=77, text_encoder_precision="fp16", tokenizer_type="clipL", reproduce=True,
                              device="cpu")
        return te, te2, byt5_model, byt5_tok, fmt

    @staticmethod
    def load_model(model_filename, model_type=None, base_model_type=None, model_def=None, quantizeTransformer=False,
                   text_encoder_quantization=None, dtype=torch.bfloat16, VAE_dtype=torch.float32, mixed_precision_transformer=False,
                   save_quantized=False, submodel_no_list=None, text_encoder_filename=None, text_encoder=None, text_encoder_2=None,
                   byt5_model=None, byt5_tokenizer=None, prompt_format=None, state_dict=None, vae_state_dict=None, vae_cfg=None,
                   vae_tiling=True, vision_encoder=None, device="cuda", **kwargs):
        """-> (pipeline_obj, pipe_dict).  Extra keyword-only hooks for tests / embedders: `state_dict` (already loaded transformer
        weights), `vae_state_dict` + `vae_cfg` (reduced VAE), `text_encoder` / `text_encoder_2` / `byt5_*` / `prompt_format` (injected
        conditioning models with the reference's protocol), `vae_tiling` (False = the one-pass B200 decode instead of the reference's tiles)."""
        from wan2gp_b200 import synth
        from wan2gp_b200.hyvideo import AutoencoderKLCausal3D, AutoencoderKLConv3D
        from wan2gp_b200.hyvideo.hunyuan import HunyuanVideoSampler
        if base_model_type not in ARCHS:
            raise ValueError(f"b200_hunyuan_handler: unsupported model type {base_model_type!r}")
        if quantizeTransformer or save_quantized:
            raise NotImplementedError("quantised transformer weights are outside the B200 bf16 hot path")
        cfg_name, v15 = ARCHS[base_model_type]
        cfg = synth.HY_CONFIGS[cfg_name]
        if state_dict is None:
            files = [model_filename] if isinstance(model_filename, str) else list(model_filename or [])
            if not files:
                raise ValueError(f"{base_model_type}: no transformer checkpoint given")
            state_dict = _read_state_dict(files[0])
        i2v = base_model_type in I2V
        model = build_transformer(cfg, v15, device, i2v=i2v)
        model.load_state_dict(state_dict)
        if i2v and "vision" not in model._g:
            raise ValueError(f"{base_model_type}: the checkpoint has no `vision_in.*` weights (a text-to-video checkpoint?)")
        model.mixed_precision = bool(mixed_precision_transformer)                  # hunyuan.py:256; selects the latent / noise dtype
        if vae_state_dict is None:
            from shared.utils import files_locator as fl                          # WanGP's checkpoint locator (hunyuan.py:326-349)
            cfg_file, sd_file = (("hunyuan_video_1_5_VAE.json", "hunyuan_video_1_5_VAE_fp32.safetensors") if v15 else
                                 ("hunyuan_video_VAE_config.json", "hunyuan_video_VAE_fp32.safetensors"))
            with open(fl.locate_file(cfg_file), "r", encoding="utf-8") as f:
                vae_cfg = json.load(f)
            vae_state_dict = _read_state_dict(fl.locate_file(sd_file))
        vae_cls = AutoencoderKLConv3D if v15 else AutoencoderKLCausal3D
        vae = vae_cls(**{k: v for k, v in dict(vae_cfg).items() if not k.startswith("_")}, device=device)
        vae.load_state_dict(vae_state_dict)
        vae._model_dtype = torch.float32 if VAE_dtype == torch.float32 else torch.bfloat16
        if text_encoder is None:
            text_encoder, text_encoder_2, byt5_model, byt5_tokenizer, prompt_format = family_handler._reference_text_encoders(
                v15, model_def, text_encoder_filename, device, text_encoder_quantization)
            if i2v and vision_encoder is None:                                   # hunyuan.py:305-309: WanGP's SigLIP encoder, in front of the path
                from models.hyvideo.vision_encoder import VisionEncoder
                from shared.utils import files_locator as fl
                vision_encoder = VisionEncoder(vision_encoder_type="siglip", vision_encoder_precision="fp16",
                                               vision_encoder_path=fl.locate_folder("siglip_vision_model"), processor_type=None, processor_path=None,
                                               output_key=None, logger=None, device=device)
                vision_encoder.vision_num_semantic_tokens, vision_encoder.vision_states_dim = 729, 1152
        pipe_obj = HunyuanVideoSampler(model, vae, text_encoder=text_encoder, text_encoder_2=text_encoder_2, byt5_model=byt5_model,
                                       byt5_tokenizer=byt5_tokenizer, prompt_format=prompt_format, hunyuan_1_5=v15, enable_cfg=v15, i2v=i2v,
                                       device=device, model_def=model_def, vae_tiling=vae_tiling, vision_encoder=vision_encoder)
        pipe = {"transformer": model, "vae": vae}
        for name, m in (("text_encoder", text_encoder), ("text_encoder_2", text_encoder_2), ("byt5_model", byt5_model),
                        ("vision_encoder", vision_encoder)):
            mod = getattr(m, "model", m)
            if isinstance(mod, torch.nn.Module):
                pipe[name] = mod
        return pipe_obj, pipe

    @staticmethod
    def fix_settings(base_model_type, settings_version, model_def, ui_defaults):
        return None

    @staticmethod
    def update_default_settings(base_model_type, model_def, ui_defaults):
        ui_defaults["embedded_guidance_scale"] = 6.0                                 # hunyuan_handler.py:302
        if base_model_type == "b200_hunyuan":
            ui_defaults.update({"guidance_scale": 7.0})                              # :304-307
        if base_model_type in I2V:
            ui_defaults.update({"image_prompt_type": "S", "sliding_window_overlap": 1})   # :342-346

    @staticmethod
    def validate_generative_settings(base_model_type, model_def, inputs):
        """Returns an error string for settings that would leave the hot path (WanGP shows it instead of queueing), else None."""
        if inputs.get("skip_steps_cache_type", "") not in ("", None):
            return "Step skipping (TeaCache / MagCache) is not available with the B200-native Hunyuan path"
        if inputs.get("activated_loras"):
            return "LoRAs are not available with the B200-native Hunyuan path"
        if base_model_type not in I2V and inputs.get("image_prompt_type", "") not in ("", None) and any(c in inputs.get("image_prompt_type", "") for c in "SVLE"):
            return "Image / video conditioning is not available with the B200-native Hunyuan text-to-video path"
        return None
