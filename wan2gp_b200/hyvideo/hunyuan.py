"""Pipeline object of the drop-in boundary, level 2 for the Hunyuan family: the B200-native counterpart of
`models/hyvideo/hunyuan.py::HunyuanVideoSampler` (:480-1085) together with the `HunyuanVideoPipeline.__call__` it drives
(`models/hyvideo/diffusion/pipelines/pipeline_hunyuan_video.py`:1100-1830), for the plain text-to-video path of Hunyuan Video 1.5
('HYVideo-1_5': CFG pair, Qwen2.5-VL + glyph-byT5 conditioning, 65 input channels of which 33 are the empty latent-concat condition)
and HunyuanVideo 1.0 ('HYVideo-T/2-cfgdistill': one forward per step with the embedded guidance scale, llava-llama-3 + CLIP-L pooled
conditioning).

WanGP's worker calls `generate(**kwargs)` (hunyuan.py:728-757 names / defaults kept verbatim) and expects `None` (aborted) or the
decoded clip `float32 CPU [3, F, H, W]` in [-1, 1] (:1080-1085: `pipeline(...)[0].squeeze(0)`); it writes `_interrupt` from the UI
thread and reports progress through `callback(step, latents, force_refresh)` (:1597 `callback(-1, None, True)`, :1763
`callback(i, latents.squeeze(0), False)`; the transformer polls `callback(-1, None, False, True)` and `_interrupt` once per block).

What runs where:
  prompt -> text encoders (INJECTED: the reference's own `TextEncoder` / `TextEncoder_1_5` / byT5 objects plug in unchanged -- the
            protocol is theirs: `text2tokens(prompts, data_type=...)`, `encode(tokens, data_type=..., device=...)` ->
            `.hidden_state`, `.attention_mask`; they sit in front of the hot path, SURVEY.md section 2 "OUT")
  -> noise with the reference's generator calls (hunyuan.py:815: one CUDA generator per sample; diffusers `randn_tensor` semantics,
     drawn in the latent dtype the reference uses -- bf16 unless `model.mixed_precision`)
  -> FlowMatchDiscreteScheduler(shift, reverse=True, solver="euler") (:864-868) x `sampling_steps` of
     `HunyuanDenoiser.step` (both CFG branches + CFG / CFG-Zero* combine + Euler update in one fused kernel)
  -> latents / scaling_factor (+ shift_factor) -> VAE decode (the reference always decodes tiled: `enable_tiling=True`, :1068).

Hunyuan 1.5 image-to-video (`i2v=True`: hunyuan_1_5_i2v) is built as well: the start frame is VAE-encoded (`latent_dist.mode() *
scaling_factor`, hunyuan.py:886-891), its latent becomes frame 0 of the 32 concat channels with mask 1 (pipeline :1513-1521) and WanGP's
SigLIP encoder's tokens enter every forward through the model's `vision_in` projection (:1528-1531, models.py:1063-1071).

Other conditioning variants of the model zoo (the token-replace i2v of HunyuanVideo 1.0, custom, custom-audio / -edit, avatar,
the 1.5 upsampler, IP / reference images, masks) are outside the hot path: asking for one raises NotImplementedError naming the
argument, which WanGP reports like any other generation error."""
import random
import re

import torch

from ..pipeline import HunyuanDenoiser
from .model import get_rotary_pos_embed

f32 = torch.float32

NEGATIVE_PROMPT = ("Aerial view, aerial view, overexposed, low quality, deformation, a poor composition, bad hands, bad teeth, bad eyes, "
                   "bad limbs, distortion")                                                       # models/hyvideo/constants.py:72
NEGATIVE_PROMPT_I2V = "deformation, a poor composition and deformed video, bad teeth, bad eyes, bad limbs"                        # :73


def align_to(value, alignment):
    """hunyuan.py / utils: round up to a multiple of `alignment`."""
    return int((value + alignment - 1) // alignment * alignment)


def randn_tensor(shape, generator, device, dtype):
    """diffusers.utils.torch_utils.randn_tensor (diffusers==0.36.0, requirements.txt) for CUDA generators: a list of one generator is
    that generator; a longer list draws sample b from generator b with batch 1 and concatenates."""
    if isinstance(generator, (list, tuple)) and len(generator) == 1:
        generator = generator[0]
    if isinstance(generator, (list, tuple)):
        return torch.cat([torch.randn((1,) + tuple(shape[1:]), generator=g, device=device, dtype=dtype) for g in generator], 0)
    return torch.randn(tuple(shape), generator=generator, device=device, dtype=dtype)


class HunyuanVideoSampler:
    """`pipeline_obj` returned by the plugin's `family_handler.load_model` for b200_hunyuan_1_5_t2v / b200_hunyuan (level 1)."""

    def __init__(self, model, vae, text_encoder=None, text_encoder_2=None, byt5_model=None, byt5_tokenizer=None, byt5_max_length=256,
                 prompt_format=None, hunyuan_1_5=True, enable_cfg=None, i2v=False, device="cuda", model_def=None, vae_tiling=True,
                 vision_encoder=None):
        if i2v and not hunyuan_1_5:
            raise NotImplementedError("HunyuanVideoSampler: the token-replace image-to-video of HunyuanVideo 1.0 is outside the hot path "
                                      "(Hunyuan 1.5 i2v = latent concat + image-encoder tokens is built)")
        self.model, self.vae = model, vae
        self.text_encoder, self.text_encoder_2 = text_encoder, text_encoder_2
        self.byt5_model, self.byt5_tokenizer, self.byt5_max_length, self.prompt_format = byt5_model, byt5_tokenizer, byt5_max_length, prompt_format
        self.hunyuan_1_5 = bool(hunyuan_1_5)
        self.enable_cfg = self.hunyuan_1_5 if enable_cfg is None else bool(enable_cfg)          # hunyuan.py:391, 413: 1.5 / custom / avatar only
        self.i2v_mode, self.custom, self.avatar, self.upsampler = bool(i2v), False, False, None
        self.vision_encoder = vision_encoder   # WanGP's SigLIP `VisionEncoder` (hunyuan.py:305-309): `.encode_images(np_uint8).last_hidden_state`
        self.device = torch.device(device)
        self.model_def = dict(model_def or {})
        self.default_negative_prompt = NEGATIVE_PROMPT_I2V if self.i2v_mode else NEGATIVE_PROMPT      # hunyuan.py:541-544
        self.vae_tiling = vae_tiling           # the reference pipeline always enables tiling before the decode (:1068, :1790-1795)
        self._interrupt = False                # written from the UI thread (wgp.py:1628); hunyuan.py:580-586 forwards it to the pipeline
        self.pipeline = self                   # `.pipeline._interrupt` is what the reference's property reads

    # ------------------------------------------------------------------ helpers
    _UNSUPPORTED = ("input_ref_images", "audio_guide", "input_frames", "input_masks", "input_video", "image_start")

    def _compression(self):
        """(spatial, temporal) compression of the attached VAE.  The reference hard-codes 16 (Hunyuan 1.5) / 8 and the "884" rule
        (hunyuan.py:683-686, 916); the production VAEs have exactly these factors, reduced test VAEs carry their own."""
        s = getattr(self.vae, "ffactor_spatial", None) or getattr(self.vae, "spatial_compression_ratio", None) or (16 if self.hunyuan_1_5 else 8)
        t = getattr(self.vae, "ffactor_temporal", None) or getattr(self.vae, "time_compression_ratio", None) or 4
        return int(s), int(t)

    def get_rotary_pos_embed(self, video_length, height, width, enable_riflex=False, spatial_compression=8, temporal_compression=4):
        """hunyuan.py:677-725 (vae "884": 4x temporal compression; theta 256; RIFLEx with L_test = latent frames)."""
        lat = [(video_length - 1) // temporal_compression + 1, height // spatial_compression, width // spatial_compression]
        ps = self.model.patch_size
        ps = [ps] * 3 if isinstance(ps, int) else list(ps)
        if any(s % p for s, p in zip(lat, ps)):
            raise AssertionError(f"Latent size(last 3 dimensions) should be divisible by patch size({ps}), but got {lat}.")
        sizes = [s // p for s, p in zip(lat, ps)]
        return get_rotary_pos_embed(sizes, self.model.rope_dim_list, theta=256.0, enable_riflex=bool(enable_riflex), L_test=lat[0])

    def _encode(self, enc, prompts, data_type, name=None, is_uncond=False):
        """encode_prompt_1_5 (:323-335, :394-400) / encode_prompt (:521-533, :617-622) through the text encoder's own protocol."""
        if self.hunyuan_1_5 and enc is self.text_encoder:
            tok = enc.text2tokens(prompts, data_type=data_type, max_length=enc.max_length)
            out = enc.encode(tok, data_type=data_type, is_uncond=True) if is_uncond else enc.encode(tok, data_type=data_type, device=self.device)
        else:
            tok = enc.text2tokens(prompts, data_type=data_type, name=name)
            out = enc.encode(tok, data_type=data_type, semantic_images=None, device=self.device)
        mask = out.attention_mask
        return out.hidden_state.to(self.device), None if mask is None else mask.to(self.device)

    def _byt5_one(self, prompt_text):
        """_process_single_byt5_prompt (:1009-1041): zeros unless the prompt quotes glyph text."""
        emb = torch.zeros((1, self.byt5_max_length, 1472), device=self.device)
        mask = torch.zeros((1, self.byt5_max_length), device=self.device, dtype=torch.int64)
        m = re.findall(r'\"(.*?)\"|“(.*?)”', prompt_text)
        glyphs = [a or b for a, b in m]
        glyphs = list(dict.fromkeys(glyphs)) if len(glyphs) > 1 else glyphs
        if glyphs and self.byt5_model is not None:
            styles = [{"color": None, "font-family": None} for _ in glyphs]
            text = self.prompt_format.format_prompt(glyphs, styles)
            tok = self.byt5_tokenizer(text, padding="max_length", max_length=self.byt5_max_length, truncation=True, add_special_tokens=True,
                                      return_tensors="pt")
            ids, mask = tok.input_ids.to(self.device), tok.attention_mask.to(self.device)
            emb = self.byt5_model(ids, attention_mask=mask.float())[0]
        return emb, mask

    # ------------------------------------------------------------------ the reference contract
    @torch.no_grad()
    def generate(self, input_prompt, input_ref_images=None, audio_guide=None, input_frames=None, input_masks=None, input_video=None, fps=24,
                 height=192, width=336, frame_num=129, seed=None, n_prompt=None, sampling_steps=50, guide_scale=1.0, shift=5.0,
                 embedded_guidance_scale=6.0, batch_size=1, num_videos_per_prompt=1, image_start=None, enable_RIFLEx=False,
                 i2v_stability=True, VAE_tile_size=None, joint_pass=False, cfg_star_switch=False, fit_into_canvas=True,
                 conditioning_latents_size=0, **kwargs):
        """Same keyword names / defaults as hunyuan.py:728-757.  -> None if interrupted, else float32 CPU [3, F, H, W] in [-1, 1]."""
        given = dict(input_ref_images=input_ref_images, audio_guide=audio_guide, input_frames=input_frames, input_masks=input_masks,
                     input_video=input_video, image_start=image_start)
        for k in self._UNSUPPORTED:
            v = given[k]
            if self.i2v_mode and k in ("input_video", "image_start"):
                continue
            if v is not None and not (isinstance(v, (list, tuple, str)) and len(v) == 0):
                raise NotImplementedError(f"HunyuanVideoSampler.generate: `{k}` belongs to a conditioning variant outside the t2v / i2v hot path")
        first_frame = None
        if self.i2v_mode:                      # hunyuan.py:879 (`first_frame = input_video[:, 0]`); a bare start image is accepted as well
            if input_video is not None:
                first_frame = input_video[:, 0]
            elif image_start is not None:
                first_frame = image_start if image_start.dim() == 3 else image_start[:, 0]
            else:
                raise ValueError("HunyuanVideoSampler.generate: image-to-video needs `input_video` (its first frame) or `image_start`")
        if VAE_tile_size is not None:                                                          # :759-772
            if self.hunyuan_1_5:
                self.vae.set_tile_sample_min_size(VAE_tile_size["tile_sample_min_size"],
                                                  VAE_tile_size.get("tile_overlap_factor", self.vae.tile_overlap_factor),
                                                  VAE_tile_size.get("tile_sample_min_tsize", self.vae.tile_sample_min_tsize))
            else:
                for k in ("tile_sample_min_tsize", "tile_latent_min_tsize", "tile_sample_min_size", "tile_latent_min_size", "tile_overlap_factor"):
                    setattr(self.vae, k, VAE_tile_size.get(k, getattr(self.vae, k)))
            self.vae.enable_tiling()
        if not self.enable_cfg:
            guide_scale = 1.0                                                                   # :775-776

        # ---- seeds (:781-815)
        n = batch_size * num_videos_per_prompt
        if isinstance(seed, torch.Tensor):
            seed = seed.tolist()
        if seed is None:
            seeds = [random.randint(0, 1_000_000) for _ in range(n)]
        elif isinstance(seed, int):
            seeds = [seed + i for _ in range(batch_size) for i in range(num_videos_per_prompt)]
        elif isinstance(seed, (list, tuple)):
            if len(seed) == batch_size:
                seeds = [int(seed[i]) + j for i in range(batch_size) for j in range(num_videos_per_prompt)]
            elif len(seed) == n:
                seeds = [int(s) for s in seed]
            else:
                raise ValueError(f"Length of seed must be equal to number of prompt(batch_size) or batch_size * num_videos_per_prompt "
                                 f"({batch_size} * {num_videos_per_prompt}), got {seed}.")
        else:
            raise ValueError(f"Seed must be an integer, a list of integers, or None, got {seed}.")
        generator = [torch.Generator(self.device).manual_seed(s) for s in seeds]

        # ---- geometry (:820-832) and prompts (:845-861)
        if width <= 0 or height <= 0 or frame_num <= 0:
            raise ValueError(f"`height` and `width` and `frame_num` must be positive integers, got height={height}, width={width}, frame_num={frame_num}")
        if (frame_num - 1) % 4 != 0:
            raise ValueError(f"`frame_num-1` must be a multiple of 4, got {frame_num}")
        target_height, target_width = align_to(height, 16), align_to(width, 16)
        if not isinstance(input_prompt, str):
            raise TypeError(f"`prompt` must be a string, but got {type(input_prompt)}")
        prompt = [input_prompt.strip()]
        if n_prompt is None or n_prompt == "":
            n_prompt = self.default_negative_prompt
        if guide_scale == 1.0:
            n_prompt = ""
        if not isinstance(n_prompt, str):
            raise TypeError(f"`negative_prompt` must be a string, but got {type(n_prompt)}")
        negative = [n_prompt.strip()]
        do_cfg = guide_scale > 1                                                                # pipeline :951-953
        comp, tcomp = self._compression()
        img_latents = vision_states = None
        if self.i2v_mode:                                                                       # hunyuan.py:886-893, pipeline :1528-1531
            ff = first_frame.to(self.device, f32)
            semantic = ff.clone().add_(1.).mul_(127.5).permute(1, 2, 0).to(torch.uint8).cpu().numpy()     # np.array(convert_tensor_to_image(frame))
            img_latents = self.vae.encode(ff[None, :, None].contiguous()).latent_dist.mode() * self.vae.config.scaling_factor   # [1, C, 1, h, w]
            target_height, target_width = int(ff.shape[1]), int(ff.shape[2])                   # `target_width, target_height = semantic_images.size`
            if self.vision_encoder is not None:
                vision_states = self.vision_encoder.encode_images(semantic).last_hidden_state.to(self.device, torch.bfloat16)
        freqs = self.get_rotary_pos_embed(frame_num, target_height, target_width, enable_RIFLEx, spatial_compression=comp,
                                          temporal_compression=tcomp)
        callback = kwargs.pop("callback", None)
        if self._interrupt:                                                                     # :1270-1271
            return None

        # ---- conditioning states (:1350-1420).  [negative, positive] order as the reference's batch.
        if self.text_encoder is None:
            raise RuntimeError("HunyuanVideoSampler: no text encoder attached (load_model wires WanGP's own TextEncoder in)")
        data_type = "video" if frame_num > 1 else "image"
        text, text_mask = self._encode(self.text_encoder, prompt, data_type)
        text_null = text_null_mask = None
        if do_cfg:
            text_null, text_null_mask = self._encode(self.text_encoder, negative, data_type, is_uncond=True)
        pooled = pooled_null = None
        if self.text_encoder_2 is not None:                                                     # CLIP-L pooled vector of HunyuanVideo 1.0
            pooled, _ = self._encode(self.text_encoder_2, prompt, data_type)
            if do_cfg:
                pooled_null, _ = self._encode(self.text_encoder_2, negative, data_type)
        byt5 = byt5_mask = byt5_null = byt5_null_mask = None
        if self.hunyuan_1_5:                                                                    # :1408-1411, :1043-1096
            byt5, byt5_mask = self._byt5_one(prompt[0])
            if do_cfg:
                byt5_null, byt5_null_mask = self._byt5_one("")
        text = text.to(torch.bfloat16)                                                          # :1469-1472
        text_null = None if text_null is None else text_null.to(torch.bfloat16)
        pooled = None if pooled is None else pooled.to(torch.bfloat16)
        pooled_null = None if pooled_null is None else pooled_null.to(torch.bfloat16)

        # ---- noise (:1493-1511 -> prepare_latents :818-900 with denoise_strength = 0: pure noise)
        lat_t = (frame_num - 1) // tcomp + 1
        C = self.model.out_channels
        shape = (n, C, lat_t, int(target_height) // comp, int(target_width) // comp)
        latent_dtype = f32 if getattr(self.model, "mixed_precision", False) else torch.bfloat16   # :1464-1467
        latents = randn_tensor(shape, generator, self.device, latent_dtype).to(f32).contiguous()
        cond_latents = None
        if self.hunyuan_1_5:       # i2v_condition_type "latent_concat" without an image: zero condition + zero mask channel (:1523-1525)
            cond_latents = torch.zeros(1, C + 1, *shape[2:], device=self.device, dtype=f32)
            if img_latents is not None:                                                         # :1513-1521: frame 0 = the image latent, mask 1 there
                if tuple(img_latents.shape[3:]) != tuple(shape[3:]) or img_latents.shape[1] != C:
                    raise ValueError(f"start-image latent {tuple(img_latents.shape)} does not match the latent geometry {shape}")
                cond_latents[:, :C, 0] = img_latents[:, :, 0].to(f32)
                cond_latents[:, C, 0] = 1.0
        guidance = None
        if embedded_guidance_scale is not None and getattr(self.model, "guidance_embed", False):
            guidance = (torch.tensor([embedded_guidance_scale], dtype=f32).to(latent_dtype) * 1000.0).to(f32)   # :1661-1670

        den = HunyuanDenoiser(self.model, num_steps=sampling_steps, shift=shift, guide_scale=guide_scale, device=self.device)
        if callback is not None:
            callback(-1, None, True)                                                            # :1597-1598

        # ---- denoising loop (:1612-1763); samples of a batch are independent: one after the other through the same kernels
        for i in range(sampling_steps):
            if self._interrupt:
                return None
            for b in range(n):
                lat = latents[b:b + 1]
                if den.step(lat, cond_latents, i, text, text_mask, text_null, text_null_mask, byt5=byt5, byt5_mask=byt5_mask, freqs=freqs,
                            text_states_2=pooled, guidance=guidance, byt5_null=byt5_null, byt5_null_mask=byt5_null_mask,
                            text_states_2_null=pooled_null, cfg_star=bool(cfg_star_switch), joint_pass=bool(joint_pass), callback=callback,
                            pipeline=self, vision_states=vision_states) is None or self._interrupt:
                    return None
            if callback is not None:
                callback(i, latents.squeeze(0), False)                                          # :1762-1763

        # ---- decode (:1777-1817)
        cfg = self.vae.config
        shift_f = getattr(cfg, "shift_factor", None)
        z = latents / cfg.scaling_factor + shift_f if shift_f else latents / cfg.scaling_factor
        if self.vae_tiling:
            self.vae.enable_tiling()
        image = self.vae.decode(z, return_dict=False, generator=generator)[0]
        if image.shape[2] == 1:
            image = image.squeeze(2)
        return image.cpu().float().squeeze(0)                                                   # :1820, hunyuan.py:1083

    # WanGP calls these around generate()
    def get_loras_transformer(self, *a, **k):
        return [], []
