"""Sampling-loop call sites of the hot path, B200-native.

Mirrors what `models/wan/any2video.py::WanAny2V.generate` does around the denoiser:
  * flow-matching Euler schedule (shared/utils/euler_scheduler.py:35-86, `sample_solver="euler"`),
  * Wan2.2 two-expert switch: the high-noise model while t > switch_threshold, then the low-noise model with its own
    guidance scale (any2video.py:1437-1443, 1491-1492; defaults/t2v_2_2.json: 875, 4 -> 3, flow_shift 12),
  * CFG: cond / uncond forwards (any2video.py:1625-1646) and the combine u + g (c - u) (:1722) fused with the
    scheduler update in one kernel,
  * VAE decode to uint8 frames (:1784).
The reference's UniPC solver, CFG-Zero*, sliding windows etc. stay in the reference (adjacent rows).
"""
import numpy as np
import torch

from . import ops

f32 = torch.float32


def euler_timesteps(num_steps, shift=5.0, num_train_timesteps=1000):
    """EulerScheduler.set_timesteps (euler_scheduler.py:35-51): linspace(1000, 1, n) ++ [0], shifted; returns n+1 values."""
    t = np.append(np.linspace(num_train_timesteps, 1, num_steps, dtype=np.float32), np.float32(0.0)).astype(np.float64)
    t = t / num_train_timesteps
    t = shift * t / (1 + (shift - 1) * t) * num_train_timesteps
    return [float(v) for v in t]


class WanDenoiser:
    """Holds the expert(s) and the schedule; `step()` is one denoise step on device-resident latents."""

    def __init__(self, model, model2=None, vae=None, num_steps=50, shift=12.0, guide_scale=4.0, guide2_scale=3.0,
                 switch_threshold=875, device="cuda", cfg_star_switch=False, cfg_zero_step=-1, cfg_group=None, cfg_rank=0):
        self.model, self.model2, self.vae = model, model2, vae
        # CFG-pair split (BASELINE configs[2]): the two ranks of `cfg_group` each run ONE branch
        # (cfg_rank 0 = cond, 1 = uncond) and exchange the fp32 prediction (19 MB at 720p x 81f) once per step; both then
        # apply the identical combine + scheduler update, so the latents stay replicated without a second collective.
        self.cfg_group, self.cfg_rank = cfg_group, cfg_rank
        self.cfg_star_switch, self.cfg_zero_step = cfg_star_switch, cfg_zero_step      # CFG-Zero* (any2video.py:1701-1722)
        self.device = torch.device(device)
        self.guide_scale, self.guide2_scale, self.switch_threshold = guide_scale, guide2_scale, switch_threshold
        self.timesteps = euler_timesteps(num_steps, shift)
        self.num_steps = num_steps
        self._interrupt = False                      # written from the UI thread in the reference (wgp.py:1628)
        self._pred = None

    def expert(self, t):
        """(model, guidance scale) for timestep t (any2video.py:1437-1443: switch when t <= switch_threshold)."""
        if self.model2 is not None and t <= self.switch_threshold:
            return self.model2, self.guide2_scale
        return self.model, self.guide_scale

    @torch.no_grad()
    def step(self, latents, i, context, context_null=None, y=None, freqs=None, callback=None):
        """latents fp32 [B,16,T,H,W] on device, updated IN PLACE; returns latents or None if interrupted."""
        t = self.timesteps[i]
        dt = (t - self.timesteps[i + 1]) / 1000.0
        model, g = self.expert(t)
        tt = torch.tensor([t], dtype=f32)
        kw = dict(y=y, freqs=freqs, pipeline=self, current_step_no=i, max_steps=self.num_steps, callback=callback)
        if context_null is None:
            cond = model([latents], tt, [context], **kw)[0]
            uncond = None
        elif self.cfg_group is not None:
            import torch.distributed as dist
            mine = model([latents], tt, [context if self.cfg_rank == 0 else context_null], **kw)[0]
            if mine is None:
                return None
            both = [torch.empty_like(mine), torch.empty_like(mine)]
            dist.all_gather(both, mine.contiguous(), group=self.cfg_group)        # ncclAllGather inside the 2-rank pair
            cond, uncond = both
        else:
            # joint pass: same blocks applied to each branch in turn (any2video.py:1634, model.py:2030-2037)
            cond, uncond = model([latents, latents], tt, [context, context_null], **kw)
        if cond is None:
            return None
        # NB any2video.py:1719 is overwritten by :1722, so steps <= cfg_zero_step are ordinary un-rescaled CFG
        self._combine_step(latents, cond, uncond, g, dt, self.cfg_star_switch and uncond is not None and i > self.cfg_zero_step)
        return latents

    @staticmethod
    def _combine_step(latents, cond, uncond, g, dt, cfg_star):
        ops.cfg_euler_step_(latents, cond, uncond, g, dt, cfg_star=cfg_star)

    @torch.no_grad()
    def step_host(self, latents_host, i, context_host, context_null_host=None, y=None, freqs=None):
        """End-to-end step with HOST buffers (pinned): H2D of the step's inputs, the step, D2H of the new latents."""
        lat = latents_host.to(self.device, non_blocking=True)
        ctx = context_host.to(self.device, non_blocking=True)
        ctxn = None if context_null_host is None else context_null_host.to(self.device, non_blocking=True)
        out = self.step(lat, i, ctx, ctxn, y=y, freqs=freqs)
        if out is None:
            return None
        latents_host.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return latents_host

    @torch.no_grad()
    def generate(self, context, context_null, latent_shape, seed=0, y=None, callback=None, decode=True):
        """Full schedule: noise -> denoise loop -> VAE decode; returns {"x": uint8 CPU [3,F,H,W]} or latents, None if aborted
        (contract of WanAny2V.generate, any2video.py:1810-1826)."""
        g = torch.Generator(device="cpu").manual_seed(seed)
        latents = torch.randn(1, *latent_shape, dtype=f32, generator=g).to(self.device)      # any2video.py:1470
        context, context_null = context.to(self.device), (None if context_null is None else context_null.to(self.device))
        for i in range(self.num_steps):
            if self.step(latents, i, context, context_null, y=y, callback=callback) is None:
                return None
        if not decode or self.vae is None:
            return {"latents": latents}
        return {"x": self.vae.decode_to_cpu_uint8([latents[0]], 0)[0]}


class HunyuanDenoiser:
    """Denoise-loop call site of the Hunyuan Video 1.5 pipeline (models/hyvideo/diffusion/pipelines/pipeline_hunyuan_video.py
    :1597-1763): latent_model_input = cat(latents, cond_latents) (:1640-1650), transformer(cond) / transformer(uncond)
    (:1655/:1687), CFG combine (:1719-1743) and the flow-matching scheduler step (:1755), here the Euler update fused with
    the combine.  guidance 6.0 / shift 9 are defaults/hunyuan_1_5_t2v.json."""

    def __init__(self, model, num_steps=30, shift=9.0, guide_scale=6.0, device="cuda"):
        self.model, self.device, self.guide_scale = model, torch.device(device), guide_scale
        self.timesteps = euler_timesteps(num_steps, shift)
        self.num_steps = num_steps
        self._interrupt = False

    @torch.no_grad()
    def step(self, latents, cond_latents, i, text, text_mask, text_null, text_null_mask, byt5=None, byt5_mask=None, freqs=None,
             text_states_2=None, guidance=None):
        """latents fp32 [1,C,T,H,W] (updated in place); cond_latents fp32 [1,C2,T,H,W] (Hunyuan 1.5 concat mask/cond channels) or
        None; text_null=None => no CFG (guidance-distilled HunyuanVideo 1.0: one forward with the guidance embedding)."""
        t = self.timesteps[i]
        dt = (t - self.timesteps[i + 1]) / 1000.0
        x = latents if cond_latents is None else torch.cat([latents, cond_latents], 1)
        tt = torch.tensor([t], dtype=f32)
        kw = dict(freqs_cos=None if freqs is None else freqs[0], freqs_sin=None if freqs is None else freqs[1], pipeline=self,
                  step_no=i, byt5_text_states=byt5, byt5_text_mask=byt5_mask, text_states_2=text_states_2, guidance=guidance)
        cond = self.model(x, tt, text_states=text, text_mask=text_mask, **kw)
        if cond is None:
            return None
        uncond = None
        if text_null is not None:
            uncond = self.model(x, tt, text_states=text_null, text_mask=text_null_mask, **kw)
            if uncond is None:
                return None
            uncond = uncond.contiguous()
        ops.cfg_euler_step_(latents, cond.contiguous(), uncond, self.guide_scale, dt)
        return latents
