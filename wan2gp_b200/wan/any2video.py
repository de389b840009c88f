"""Pipeline object of the drop-in boundary, level 2: the B200-native counterpart of
`models/wan/any2video.py::WanAny2V` for the plain t2v / i2v2_2 generation path.

WanGP's worker thread calls `wan_model.generate(**kwargs)` with ~110 keyword arguments (wgp.py:7762-7880) and expects
`None` (aborted) or `{"x": uint8 CPU tensor [3,F,H,W], "latent_slice": ...}` (any2video.py:1810-1826; consumed at wgp.py:7937-7966);
it writes `wan_model._interrupt` from the UI thread, reads `.model`, `.model2`, `.vae`, and reports progress through
`callback(step, latents_preview, force_refresh, **kw)`.  This class keeps that contract -- same keyword names and defaults
(any2video.py:414-503), same call protocol of `callback`, same exceptions-as-errors behaviour -- around the B200 hot path:

  text encoder (injected; the umT5 encoder is the reference's own, SURVEY.md section 8 puts it before the path)
  -> [i2v2_2: WanVAE.encode of the start image / prefix video + 4-channel frame mask, any2video.py:650-775]
  -> noise (same CUDA-generator call as any2video.py:1470) -> WanDenoiser.step x N (fused CFG / CFG-Zero* + scheduler kernel)
     with the reference's guidance-phase / expert switch (update_guidance, any2video.py:1437-1443)
  -> WanVAE.decode_to_cpu_uint8 (any2video.py:1784).

Conditioning variants of the model zoo (VACE, multitalk, phantom, ... ), NAG, APG, skip-layer guidance, sliding-window overlap
latents and the self-refiner are out of scope of the hot path: asking for one raises NotImplementedError naming it, which WanGP
reports like any other generation error (wgp.py:7883-7919)."""
import torch

from ..pipeline import WanDenoiser
from .rope import get_rotary_pos_embed

f32 = torch.float32

SAMPLE_NEG_PROMPT = ("色调艳丽，过曝，静态，细节模糊不清，字幕，风格，作品，画作，画面，静止，整体发灰，最差质量，低质量，JPEG压缩残留，丑陋的，残缺的，多余的手指，"
                     "画得不好的手部，画得不好的脸部，畸形的，毁容的，形态畸形的肢体，手指融合，静止不动的画面，杂乱的背景，三条腿，背景人很多，倒着走")   # shared_config.py


def _is_set(v):
    if v is None:
        return False
    if isinstance(v, (list, tuple, dict, str)):
        return len(v) > 0
    return True


class WanAny2V:
    """`pipeline_obj` returned by `family_handler.load_model` (level 1) for base model types t2v / t2v_1.3B / t2v_2_2 / i2v_2_2."""

    vae_stride = (4, 8, 8)
    patch_size = (1, 2, 2)
    num_train_timesteps = 1000

    def __init__(self, model, model2=None, vae=None, text_encoder=None, model_def=None, base_model_type="t2v_2_2", device="cuda",
                 dtype=torch.bfloat16, VAE_dtype=torch.float32):
        self.model, self.model2, self.vae, self.vae2 = model, model2, vae, None
        self.text_encoder = text_encoder                       # callable(prompts: list[str], device) -> list[Tensor[len, text_dim]]
        self.model_def = dict(model_def or {})
        self.base_model_type = base_model_type
        self.device, self.dtype, self.VAE_dtype = torch.device(device), dtype, VAE_dtype
        self.sample_neg_prompt = SAMPLE_NEG_PROMPT
        self._interrupt = False                                 # written from the UI thread (wgp.py:1628)
        self.i2v = bool(self.model_def.get("i2v_2_2", False)) or getattr(model, "model_type", "") == "i2v2_2"

    # ------------------------------------------------------------------ helpers
    _UNSUPPORTED = {   # kwarg -> the value that means "not requested"
        "input_frames": None, "input_frames2": None, "input_masks": None, "input_masks2": None, "input_ref_images": None,
        "input_ref_masks": None, "input_faces": None, "input_custom": None, "image_end": None, "target_camera": None,
        "audio_proj": None, "audio_scale": None, "overlapped_latents": None, "face_arc_embeds": None, "vae_upsampler": None,
        "perturbation_layers": None, "speakers_bboxes": None, "prefix_video": None,
    }

    def _check_scope(self, kw):
        for k in self._UNSUPPORTED:
            if _is_set(kw.get(k)):
                raise NotImplementedError(f"WanAny2V.generate: `{k}` belongs to a conditioning variant outside the t2v / i2v2_2 hot path")
        if kw.get("NAG_scale", 0) > 1:
            raise NotImplementedError("NAG guidance is outside the hot path")
        if kw.get("apg_switch", False):
            raise NotImplementedError("adaptive projected guidance is outside the hot path")
        if kw.get("self_refiner_setting", 0) > 0:
            raise NotImplementedError("the self-refiner is outside the hot path")
        if kw.get("sub_parallel_window_size", 0) > 0:
            raise NotImplementedError("sub-parallel windows are outside the hot path")
        if "G" in (kw.get("video_prompt_type") or ""):     # denoising_strength / masking_strength only act in this mode (any2video.py:1215-1240)
            raise NotImplementedError("video-to-video (denoising / masking strength) is outside the hot path")
        if kw.get("image_mode", 0) != 0:
            raise NotImplementedError("image outputs (image_mode) are outside the hot path")

    def _encode_prompt(self, prompt, text_len):
        """any2video.py:590-595: encode, cast, zero-pad to text_len, add the batch dim."""
        if self.text_encoder is None:
            raise RuntimeError("WanAny2V: no text encoder attached (load_model wires the reference's umT5 encoder in)")
        ctx = self.text_encoder([prompt], self.device)[0].to(self.device, f32)
        if ctx.shape[0] > text_len:
            ctx = ctx[:text_len]
        return torch.cat([ctx, ctx.new_zeros(text_len - ctx.shape[0], ctx.shape[1])]).unsqueeze(0)

    def _i2v_condition(self, image_start, input_video, frame_num, height, width, VAE_tile_size, motion_amplitude):
        """y = cat(mask [4, lat_T, h, w], VAE.encode([start frames ++ zeros]) [16, lat_T, h, w]) -- any2video.py:667-775 for the
        plain i2v2_2 case (no end frame, no svi / infinitetalk branches)."""
        if input_video is None:
            if image_start is None:
                input_video = torch.full((3, 1, height, width), -1.0)              # :667-669
            else:
                input_video = image_start if image_start.dim() == 4 else image_start.unsqueeze(1)
        _, pre, height, width = input_video.shape
        input_video = input_video.to(self.device, f32)
        lat_h, lat_w = height // self.vae_stride[1], width // self.vae_stride[2]
        enc = torch.cat([input_video, torch.zeros(3, frame_num - pre, height, width, device=self.device, dtype=f32)], 1)
        lat_y = self.vae.encode([enc], VAE_tile_size)[0]
        msk = torch.ones(1, frame_num, lat_h, lat_w, device=self.device)
        msk[:, pre:] = 0
        msk = torch.cat([torch.repeat_interleave(msk[:, 0:1], repeats=4, dim=1), msk[:, 1:]], 1)
        msk = msk.view(1, msk.shape[1] // 4, 4, lat_h, lat_w).transpose(1, 2)[0]
        if motion_amplitude > 1:                                                   # :760-771
            base = lat_y[:, :1]
            diff = lat_y[:, pre:] - base
            mean = diff.mean(dim=(0, 2, 3), keepdim=True)
            lat_y = torch.cat([lat_y[:, :pre], torch.clamp(base + (diff - mean) * motion_amplitude + mean, -6, 6)], 1)
        return torch.cat([msk, lat_y.to(self.device, f32)]), height, width

    # ------------------------------------------------------------------ the reference contract
    @torch.no_grad()
    def generate(self, input_prompt, alt_prompt="", input_frames=None, input_frames2=None, input_masks=None, input_masks2=None,
                 input_ref_images=None, input_ref_masks=None, input_faces=None, input_video=None, image_start=None, image_end=None,
                 input_custom=None, denoising_strength=1.0, masking_strength=1.0, target_camera=None, context_scale=None, width=1280,
                 height=720, fit_into_canvas=True, frame_num=81, batch_size=1, shift=5.0, sample_solver="unipc", sampling_steps=50,
                 guide_scale=5.0, guide2_scale=5.0, guide3_scale=5.0, switch_threshold=0, switch2_threshold=0, guide_phases=1,
                 model_switch_phase=1, n_prompt="", seed=-1, callback=None, enable_RIFLEx=None, VAE_tile_size=0, joint_pass=False,
                 perturbation_layers=None, perturbation_start=0.0, perturbation_end=1.0, cfg_star_switch=True, cfg_zero_step=5,
                 audio_scale=None, audio_cfg_scale=None, audio_proj=None, audio_context_lens=None, alt_guide_scale=1.0,
                 overlapped_latents=None, return_latent_slice=None, overlap_noise=0, overlap_size=0, sub_parallel_window_size=0,
                 sub_parallel_window_overlap=0, conditioning_latents_size=0, keep_frames_parsed=[], model_type=None, model_mode=None,
                 loras_slists=None, NAG_scale=0, NAG_tau=3.5, NAG_alpha=0.5, offloadobj=None, apg_switch=False, speakers_bboxes=None,
                 color_correction_strength=1, prefix_frames_count=0, image_mode=0, window_no=0, set_header_text=None,
                 pre_video_frame=None, prefix_video=None, video_prompt_type="", original_input_ref_images=[], face_arc_embeds=None,
                 control_scale_alt=1., motion_amplitude=1., window_start_frame_no=0, self_refiner_setting=0, self_refiner_plan="",
                 self_refiner_f_uncertainty=0.0, self_refiner_certain_percentage=0.999, custom_settings=None, save_masks=False,
                 vae_upsampler=None, set_progress_status=None, fps=16, **bbargs):
        """Same keyword names / defaults as any2video.py:414-503.  Returns None if interrupted, else {"x": uint8 [3,F,H,W] on the CPU,
        "latent_slice": Tensor | None}."""
        self._check_scope(dict(locals(), **bbargs))
        if self._interrupt:
            return None
        if seed is None or seed < 0:
            seed = int(torch.seed() % (2 ** 31))
        if n_prompt == "":
            n_prompt = self.sample_neg_prompt
        text_len = self.model.text_len
        any_guidance_at_all = guide_scale > 1 or (guide2_scale > 1 and guide_phases >= 2) or (guide3_scale > 1 and guide_phases >= 3)   # :571
        context = self._encode_prompt(input_prompt, text_len)
        context_null = self._encode_prompt(n_prompt, text_len) if any_guidance_at_all else None
        if offloadobj is not None and hasattr(offloadobj, "unload_all"):
            offloadobj.unload_all()                                                 # :601 (mmgp; nothing is offloaded on a B200)
        if self._interrupt:
            return None

        # ---- conditioning (i2v2_2) and latent geometry -- :647, :1166
        y = None
        if self.i2v:
            y, height, width = self._i2v_condition(image_start, input_video, frame_num, height, width, VAE_tile_size, motion_amplitude)
        elif _is_set(image_start) or input_video is not None:
            raise NotImplementedError("image / video conditioning needs an i2v_2_2 model definition")
        lat_frames = int((frame_num - 1) // self.vae_stride[0]) + 1
        lat_h, lat_w = height // self.vae_stride[1], width // self.vae_stride[2]
        z_dim = getattr(getattr(self.vae, "model", None), "z_dim", 16) if self.vae is not None else 16
        target_shape = (z_dim, lat_frames, lat_h, lat_w)
        freqs = get_rotary_pos_embed(target_shape[1:], enable_RIFLEx=bool(enable_RIFLEx))                       # :1192

        # ---- scheduler + guidance phases -- :506-545, :1425-1443
        den = WanDenoiser(self.model, self.model2, self.vae, num_steps=sampling_steps, shift=shift, guide_scale=guide_scale,
                          guide2_scale=guide2_scale, switch_threshold=switch_threshold, device=self.device,
                          cfg_star_switch=bool(cfg_star_switch), cfg_zero_step=cfg_zero_step, sample_solver=sample_solver)
        den.interrupt_source = self
        timesteps = den.timesteps[:-1]
        state = {"guide": guide_scale, "trans": self.model, "done2": False, "done3": False, "extra": ""}
        if guide_phases > 1:
            state["extra"] = f"Phase 1/{guide_phases} High Noise" if self.model2 is not None else f"Phase 1/{guide_phases}"

        def update_guidance(step_no, t, new_guide, done_key, threshold, phase_no):
            """any2video.py:1437-1443."""
            if guide_phases >= phase_no and not state[done_key] and t <= threshold:
                if model_switch_phase == phase_no - 1 and self.model2 is not None:
                    state["trans"] = self.model2
                state["guide"], state[done_key] = new_guide, True
                low = state["trans"] is self.model2
                state["extra"] = (f"Phase {phase_no}/{guide_phases} {'Low Noise' if low else 'High Noise'}" if self.model2 is not None
                                  else f"Phase {phase_no}/{guide_phases}")
                if callback is not None:
                    callback(step_no - 1, denoising_extra=state["extra"])

        den.expert = lambda t: (state["trans"], state["guide"])                    # the phase logic above replaces the fixed t <= threshold rule
        if callback is not None:
            callback(-1, None, True)                                                # :1409
            callback(-1, None, True, override_num_inference_steps=len(timesteps), denoising_extra=state["extra"])   # :1446

        # ---- noise: same generator call as the reference (:548-549, :1470) so a seed reproduces the reference's latents on a GPU
        seed_g = torch.Generator(device=self.device)
        seed_g.manual_seed(seed)
        latents = torch.randn(batch_size, *target_shape, dtype=f32, device=self.device, generator=seed_g)

        # ---- denoising loop -- :1490-1750
        for i, t in enumerate(timesteps):
            update_guidance(i, t, guide2_scale, "done2", switch_threshold, 2)
            update_guidance(i, t, guide3_scale, "done3", switch2_threshold, 3)
            null = context_null if (any_guidance_at_all and state["guide"] != 1) else None     # :1606 any_guidance = guide_scale != 1
            if den.step(latents, i, context, null, y=y, freqs=freqs, callback=callback) is None or self._interrupt:
                return None
            if callback is not None:
                preview = latents if latents.shape[0] == 1 else latents.transpose(0, 2)          # :1741-1746
                callback(i, preview[0], False, denoising_extra=state["extra"])

        latent_slice = latents[:, :, return_latent_slice].clone() if return_latent_slice is not None else None   # :1757-1758
        videos = self.vae.decode_to_cpu_uint8(list(latents.unbind(0)), VAE_tile_size)           # :1784
        return {"x": videos[0], "latent_slice": latent_slice}                                  # :1803, :1810 (first video only)

    # the reference exposes these on the pipeline object; WanGP calls them around generate()
    def get_loras_transformer(self, *a, **k):
        return [], []

    def get_trans_lora(self, *a, **k):
        return self.model, None
