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

from . import _lib, ops

f32 = torch.float32


def euler_timesteps(num_steps, shift=5.0, num_train_timesteps=1000):
    """EulerScheduler.set_timesteps (euler_scheduler.py:35-51): linspace(1000, 1, n) ++ [0], shifted; returns n+1 values."""
    t = np.append(np.linspace(num_train_timesteps, 1, num_steps, dtype=np.float32), np.float32(0.0)).astype(np.float64)
    t = t / num_train_timesteps
    t = shift * t / (1 + (shift - 1) * t) * num_train_timesteps
    return [float(v) for v in t]


def flow_match_timesteps(num_steps, shift=7.0, num_train_timesteps=1000):
    """FlowMatchDiscreteScheduler.set_timesteps with reverse=True, solver="euler" (models/hyvideo/diffusion/schedulers/
    scheduling_flow_match_discrete.py:123-151, 183-184; built at hunyuan.py:864-868): sigmas = shift(linspace(1, 0, n+1)) in fp32; returns
    the n+1 values * 1000 (the last is 0).  NB the grid differs from the Wan EulerScheduler's (linspace(1000, 1, n) ++ [0])."""
    sig = np.linspace(1, 0, num_steps + 1, dtype=np.float32)
    sig = (np.float32(shift) * sig / (1 + (np.float32(shift) - 1) * sig)).astype(np.float32)
    return [float(v) for v in sig.astype(np.float64) * num_train_timesteps]


def lcm_timesteps(num_steps, shift=5.0, num_train_timesteps=1000):
    """LCMScheduler.set_timesteps (shared/utils/lcm_scheduler.py:26-57; `sample_solver="lcm"`, any2video.py:533-543): at most 8 steps, sigmas
    from 1 down to sigma_min = 0.003/1.002 (NOT 0), shifted; its step (:59-87) is the Euler update x += v (sigma_next - sigma)."""
    n = min(int(num_steps), 8)
    t = np.linspace(0, 1, n + 1, dtype=np.float32)
    smin = np.float32(0.003 / 1.002)
    sig = (smin + (np.float32(1.0) - smin) * (1 - t)).astype(np.float32)
    sig = (np.float32(shift) * sig / (1 + (np.float32(shift) - 1) * sig)).astype(np.float32)
    return [float(v) for v in sig.astype(np.float64) * num_train_timesteps]


def causvid_timesteps(num_steps):
    """`sample_solver="causvid"` (any2video.py:513-517): fixed table, sigmas = t/1000 ++ [0], FlowMatchScheduler.step (basic_flowmatch.py:45-57)
    = Euler."""
    return [float(v) for v in [1000, 934, 862, 756, 603, 410, 250, 140, 74][:num_steps]] + [0.0]


class UniPCSchedule:
    """Host side of FlowUniPCMultistepScheduler (shared/utils/fm_solvers_unipc.py), WanGP's default `sample_solver="unipc"`
    (any2video.py:518-522): solver_order 2, bh2, predict_x0, flow_prediction, lower_order_final, corrector on every step > 0.
    Every tensor update of one `step()` is a linear combination of {x, v, last_sample, x0_{i-1}, x0_{i-2}}; this class computes
    the scalar coefficients (float64) and `ops.cfg_unipc_step_` applies them in ONE kernel fused with the CFG combine:
        x0 = x - sigma_i v                                                   convert_model_output (:300-333)
        xc = ca x_last + cb m0 + cc m1 + cd x0      (steps > 0, else xc = x)  multistep_uni_c_bh_update (:500-640)
        xn = pp xc + pq x0 + pr m0                                            multistep_uni_p_bh_update (:345-470)
    with m0 = x0_{i-1}, m1 = x0_{i-2}."""

    def __init__(self, num_steps, shift=5.0, num_train_timesteps=1000):
        # __init__ (:108-138): sigma_max/min of the training schedule with shift 1; set_timesteps (:158-225)
        train = 1.0 - np.linspace(1, 1 / num_train_timesteps, num_train_timesteps)[::-1]
        sig = np.linspace(float(np.float32(train[0])), float(np.float32(train[-1])), num_steps + 1)[:-1]
        sig = shift * sig / (1 + (shift - 1) * sig)
        self.timesteps = [int(v) for v in (sig * num_train_timesteps).astype(np.int64)]      # int64 truncation, as the reference
        self.sigmas = np.concatenate([sig, [0.0]]).astype(np.float32).astype(np.float64)
        self.num_steps, self.solver_order = num_steps, 2
        self.reset()

    def reset(self):
        self.lower_order_nums, self.this_order = 0, 1

    @staticmethod
    def _lam(sigma):
        with np.errstate(divide="ignore"):
            return np.log(1.0 - sigma) - np.log(sigma)

    def _bh(self, i_t, i_s0, i_prev, order):
        """Common part of UniP / UniC for the update from sigma[i_s0] to sigma[i_t]: (sigma ratio, alpha_t*h_phi_1, alpha_t*B_h, rk,
        R, b) -- :392-447 / :571-612."""
        sg = self.sigmas
        sigma_t, sigma_s0 = sg[i_t], sg[i_s0]
        alpha_t = 1.0 - sigma_t
        h = self._lam(sigma_t) - self._lam(sigma_s0)
        hh = -h
        with np.errstate(invalid="ignore", over="ignore"):
            h_phi_1 = np.expm1(hh)
            B_h = np.expm1(hh)
            rks = []
            if order == 2:
                rks.append((self._lam(sg[i_prev]) - self._lam(sigma_s0)) / h)
            rks.append(1.0)
            rks = np.array(rks)
            R, b, h_phi_k, fact = [], [], h_phi_1 / hh - 1, 1
            for k in range(1, order + 1):
                R.append(rks ** (k - 1))
                b.append(h_phi_k * fact / B_h)
                fact *= k + 1
                h_phi_k = h_phi_k / hh - 1 / fact
        return sigma_t / sigma_s0, alpha_t * h_phi_1, alpha_t * B_h, rks[0], np.stack(R), np.array(b)

    def coefficients(self, i):
        """Scalars of step i (call in order i = 0, 1, ...): dict(sigma, use_corrector, ca, cb, cc, cd, pp, pq, pr)."""
        c = dict(sigma=float(self.sigmas[i]), use_corrector=i > 0, ca=0.0, cb=0.0, cc=0.0, cd=0.0)
        if i > 0:                                                     # UniC over sigma[i-1] -> sigma[i] with the PREVIOUS step's order
            order = self.this_order
            ratio, a_phi, a_bh, rk, R, b = self._bh(i, i - 1, i - 2, order)
            rhos = np.array([0.5]) if order == 1 else np.linalg.solve(R, b)
            k1 = rhos[0] / rk if order == 2 else 0.0
            c.update(ca=ratio, cb=-a_phi + a_bh * (k1 + rhos[-1]), cc=-a_bh * k1, cd=-a_bh * rhos[-1])
        # order of this step's predictor (:707-716): lower_order_final + multistep warm-up
        self.this_order = min(min(self.solver_order, self.num_steps - i), self.lower_order_nums + 1)
        ratio, a_phi, a_bh, rk, _, _ = self._bh(i + 1, i, i - 1, self.this_order)
        k1 = 0.5 / rk if self.this_order == 2 else 0.0                # rhos_p = [0.5] for order 2 (:453-455)
        c.update(pp=ratio, pq=-a_phi + a_bh * k1, pr=-a_bh * k1)
        if self.lower_order_nums < self.solver_order:
            self.lower_order_nums += 1
        return {k: (v if isinstance(v, bool) else float(v)) for k, v in c.items()}


class DPMppSchedule:
    """Host side of FlowDPMSolverMultistepScheduler (shared/utils/fm_solvers.py; `sample_solver="dpm++"`, any2video.py:523-532):
    dpmsolver++ order 2, midpoint, final_sigmas_type "zero" (last step first order), sigmas = get_sampling_sigmas(steps, shift).
    Its update is the same linear form the UniPC kernel applies, without a corrector (fm_solvers.py:459-478, 529-560):
        x0 = x - sigma_i v;   xn = pp x + pq x0 + pr x0_{i-1}
    first order:  pp = sigma_t/sigma_s, pq = -alpha_t (e^{-h} - 1);   second:  pq *= (1 + 1/(2 r0)), pr = alpha_t (e^{-h} - 1) / (2 r0)."""

    def __init__(self, num_steps, shift=5.0, num_train_timesteps=1000):
        sig = np.linspace(1, 0, num_steps + 1)[:num_steps]                       # get_sampling_sigmas (fm_solvers.py:22-26)
        sig = shift * sig / (1 + (shift - 1) * sig)
        self.timesteps = [int(v) for v in (sig * num_train_timesteps).astype(np.int64)]
        self.sigmas = np.concatenate([sig, [0.0]]).astype(np.float32).astype(np.float64)
        self.num_steps = num_steps
        self.reset()

    def reset(self):
        self.lower_order_nums = 0

    def coefficients(self, i):
        sg, lam = self.sigmas, UniPCSchedule._lam
        sigma_t, sigma_s0 = sg[i + 1], sg[i]
        alpha_t = 1.0 - sigma_t
        with np.errstate(invalid="ignore", over="ignore", divide="ignore"):
            h = lam(sigma_t) - lam(sigma_s0)
            coef = -alpha_t * np.expm1(-h)
            first = self.lower_order_nums < 1 or i == self.num_steps - 1         # lower_order_final with final sigma zero (:745-749)
            k1 = 0.0 if first else 0.5 / ((lam(sigma_s0) - lam(sg[i - 1])) / h)
        c = dict(sigma=float(sg[i]), use_corrector=False, ca=0.0, cb=0.0, cc=0.0, cd=0.0,
                 pp=float(sigma_t / sigma_s0), pq=float(coef * (1 + k1)), pr=float(-coef * k1))
        if self.lower_order_nums < 2:
            self.lower_order_nums += 1
        return c


def exchange_cfg_pair(mine, like, group):
    """The one per-step exchange of a CFG pair split over two ranks: every rank contributes its branch's fp32
    prediction (`mine`, or None if its forward was interrupted) and gets (prediction of pair rank 0, prediction of pair rank 1), or None
    if EITHER rank was interrupted.  `_interrupt` is per process: the pair must agree on the abort before either rank skips the
    collective, otherwise the partner blocks in ncclAllGather forever -- the flag rides in front of the prediction (same collective,
    16 bytes: the predictions behind the header stay aligned for the 128-bit loads of the step kernels)."""
    import torch.distributed as dist
    HDR = 4
    shape = tuple(like.shape)
    payload = torch.zeros(HDR + like.numel(), device=like.device, dtype=f32)
    if mine is None:
        payload[0] = 1.0
    else:
        payload[HDR:].copy_(mine.reshape(-1))
    both = [torch.empty_like(payload), torch.empty_like(payload)]
    dist.all_gather(both, payload, group=group)                                   # ncclAllGather inside the 2-rank pair
    if float(both[0][0]) + float(both[1][0]) > 0:
        return None
    return both[0][HDR:].reshape(shape), both[1][HDR:].reshape(shape)


class WanDenoiser:
    """Holds the expert(s) and the schedule; `step()` is one denoise step on device-resident latents."""

    def __init__(self, model, model2=None, vae=None, num_steps=50, shift=12.0, guide_scale=4.0, guide2_scale=3.0,
                 switch_threshold=875, device="cuda", cfg_star_switch=False, cfg_zero_step=-1, cfg_group=None, cfg_rank=0,
                 sample_solver="euler"):
        self.model, self.model2, self.vae = model, model2, vae
        for m in (model, model2):
            if m is not None and hasattr(m, "cache_context"):
                m.cache_context = True           # prompts are fixed for the whole schedule: project the text once
        # CFG-pair split (BASELINE configs[2]): the two ranks of `cfg_group` each run ONE branch
        # (cfg_rank 0 = cond, 1 = uncond) and exchange the fp32 prediction (19 MB at 720p x 81f) once per step; both then
        # apply the identical combine + scheduler update, so the latents stay replicated without a second collective.
        self.cfg_group, self.cfg_rank = cfg_group, cfg_rank
        self.cfg_star_switch, self.cfg_zero_step = cfg_star_switch, cfg_zero_step      # CFG-Zero* (any2video.py:1701-1722)
        self.device = torch.device(device)
        self.guide_scale, self.guide2_scale, self.switch_threshold = guide_scale, guide2_scale, switch_threshold
        if sample_solver not in ("euler", "unipc", "", "dpm++", "lcm", "causvid"):
            raise NotImplementedError(f"sample_solver {sample_solver!r}: euler, unipc (the WanGP default), dpm++, lcm and causvid are built")
        multistep = {"unipc": UniPCSchedule, "": UniPCSchedule, "dpm++": DPMppSchedule}.get(sample_solver)
        self.unipc = None if multistep is None else multistep(num_steps, shift)
        if self.unipc is not None:
            self.timesteps = [float(t) for t in self.unipc.timesteps] + [0.0]
        else:                                     # single-step solvers: one Euler kernel, different sigma tables
            self.timesteps = {"euler": lambda: euler_timesteps(num_steps, shift), "lcm": lambda: lcm_timesteps(num_steps, shift),
                              "causvid": lambda: causvid_timesteps(num_steps)}[sample_solver]()
        num_steps = len(self.timesteps) - 1
        self.num_steps = num_steps
        self._interrupt = False                      # written from the UI thread in the reference (wgp.py:1628)
        self.interrupt_source = None                 # pipeline object that owns `_interrupt` when this denoiser is driven by WanAny2V
        self._pred, self._hist = None, None
        # whole-step CUDA graph: both CFG forwards, the combine and the Euler update of one step captured
        # once per (expert, latent buffer, prompt, CFG-Zero* phase) and replayed with the timestep / guidance / dt read from device
        # memory.  For launch-bound configurations (the 1.3B model: ~400 launches of 10-100 us per step); the reference's per-block
        # interrupt poll becomes a per-step poll.  Every solver (Euler-type and the multi-step UniPC / dpm++: one capture per step
        # parity, the history buffers swap roles); the CFG-pair split over two GPUs takes the ordinary path.
        self.use_step_graph = False
        self.graph_launches = 0
        self._step_graphs, self._ghist = {}, None
        self._staging = {}

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
        if self.use_step_graph and self.cfg_group is None and context_null is not None:
            star = self.cfg_star_switch and i > self.cfg_zero_step
            return self._graph_step(model, g, latents, i, t, dt, context, context_null, y, freqs, star)
        tt = torch.tensor([t], dtype=f32)
        # the model polls `pipeline._interrupt` once per block: the object the UI thread writes to (WanAny2V) when one is attached
        kw = dict(y=y, freqs=freqs, pipeline=self.interrupt_source or self, current_step_no=i, max_steps=self.num_steps, callback=callback)
        if context_null is None:
            cond = model([latents], tt, [context], **kw)[0]
            uncond = None
        elif self.cfg_group is not None:
            mine = model([latents], tt, [context if self.cfg_rank == 0 else context_null], **kw)[0]
            pair = exchange_cfg_pair(mine, latents, self.cfg_group)
            if pair is None:
                self._interrupt = True                                            # both ranks leave the schedule together
                return None
            cond, uncond = pair
        else:
            # joint pass: same blocks applied to each branch in turn (any2video.py:1634, model.py:2030-2037)
            cond, uncond = model([latents, latents], tt, [context, context_null], **kw)
        if cond is None:
            return None
        # NB any2video.py:1719 is overwritten by :1722, so steps <= cfg_zero_step are ordinary un-rescaled CFG
        star = self.cfg_star_switch and uncond is not None and i > self.cfg_zero_step
        if self.unipc is None:
            self._combine_step(latents, cond, uncond, g, dt, star)
        else:
            if i == 0 or self._hist is None or self._hist[0].shape != latents.shape:
                self.unipc.reset()
                self._hist = [torch.zeros_like(latents) for _ in range(3)]          # last_sample, x0_{i-1}, x0_{i-2}
            x_last, m0, m1 = self._hist
            ops.cfg_unipc_step_(latents, cond, uncond, g, x_last, m0, m1, self.unipc.coefficients(i), cfg_star=star)
            self._hist = [x_last, m1, m0]                                           # the kernel stored x0_i into m1
        return latents

    def _graph_step(self, model, g, latents, i, t, dt, context, context_null, y, freqs, star):
        if (self.interrupt_source or self)._interrupt:
            return None
        multistep = self.unipc is not None
        hist = None
        if multistep:
            # UniPC / dpm++: host coefficients of step i (stateful, call in order) -> 10 floats in device memory; the history buffers
            # have fixed addresses and swap roles every step, so there is one capture per step parity
            if i == 0 or self._ghist is None or self._ghist[0].shape != latents.shape:
                self.unipc.reset()
                if self._ghist is None or self._ghist[0].shape != latents.shape:
                    self._ghist = [torch.zeros_like(latents) for _ in range(3)]            # last_sample, x0 of the two previous steps
                else:
                    for h in self._ghist:
                        h.zero_()
            c = self.unipc.coefficients(i)
            params = [float(g)] + [float(c[k]) for k in ("sigma", "ca", "cb", "cc", "cd", "pp", "pq", "pr")] + [float(bool(c["use_corrector"]))]
            hist = (self._ghist[0],) + ((self._ghist[1], self._ghist[2]) if i % 2 == 0 else (self._ghist[2], self._ghist[1]))
        else:
            params = [float(g), float(dt)]
        # operands are identified like the model's prompt cache does (address, version, shape): a captured graph replays the text
        # projections cached at capture time, so a prompt tensor that was written to since must not hit; the entry keeps the
        # tensors alive so that their addresses cannot be recycled
        ident = lambda x: None if x is None else (x.data_ptr(), x._version, tuple(x.shape))
        key = (id(model), getattr(model, "weights_version", 0), latents.data_ptr(), tuple(latents.shape), ident(context), ident(context_null),
               ident(y), None if freqs is None else (ident(freqs[0]), ident(freqs[1])), bool(star), multistep, (i & 1) if multistep else 0)
        ent = self._step_graphs.get(key)
        if ent is None:
            tdev = torch.zeros(1, device=latents.device, dtype=f32)
            pdev = torch.zeros(len(params), device=latents.device, dtype=f32)

            class _Quiet:                                   # nothing can interrupt a capture; the poll moves to the step boundary
                _interrupt = False
            prev = getattr(model, "use_cuda_graphs", False)
            model.use_cuda_graphs = False                   # per-block graphs cannot be replayed inside a capture

            def body(lat, hs):
                cond, uncond = model([lat, lat], tdev, [context, context_null], y=y, freqs=freqs, pipeline=_Quiet(), current_step_no=0,
                                     max_steps=self.num_steps, callback=None)
                if multistep:
                    ops.cfg_unipc_step_(lat, cond, uncond, 0.0, hs[0], hs[1], hs[2], pdev, cfg_star=star)
                else:
                    ops.cfg_euler_step_(lat, cond, uncond, 0.0, pdev, cfg_star=star)
            # one eager step on SCRATCH copies of the latents (and of the solver history) fills every cache (RoPE tables, text
            # projections, cross K/V, kernel attributes) with the real values of this step: host-to-device copies and first-use
            # attribute calls are not capturable
            tdev.fill_(float(t))
            pdev.copy_(torch.tensor(params, dtype=f32))
            body(latents.clone(), None if hist is None else tuple(h.clone() for h in hist))
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            n0 = _lib.launch_count()
            with torch.cuda.graph(graph):
                body(latents, hist)
            n_kernels = _lib.launch_count() - n0            # kernels of ours recorded in the graph (the C ABI counts at launch = capture time)
            model.use_cuda_graphs = prev
            ent = self._step_graphs[key] = (graph, tdev, pdev, (latents, context, context_null, y, hist, freqs), n_kernels)   # freqs: RIFLEx tables differ at equal shapes
            if len(self._step_graphs) > 12:                 # each entry pins a memory pool: keep the table small
                self._step_graphs.pop(next(iter(self._step_graphs)))
        graph, tdev, pdev = ent[:3]
        tdev.fill_(float(t))                                # one scalar fill, one small copy and one graph launch per step
        pdev.copy_(torch.tensor(params, dtype=f32))
        graph.replay()
        self.graph_launches += ent[4]                       # kernels launched through graph replays (b200_launch_count does not see them)
        return latents

    @staticmethod
    def _combine_step(latents, cond, uncond, g, dt, cfg_star):
        ops.cfg_euler_step_(latents, cond, uncond, g, dt, cfg_star=cfg_star)

    @torch.no_grad()
    def step_host(self, latents_host, i, context_host, context_null_host=None, y=None, freqs=None):
        """End-to-end step with HOST buffers (pinned): H2D of the step's inputs, the step, D2H of the new latents."""
        def stage(name, host):                       # stable device staging buffers: same addresses every step (whole-step graph key)
            if host is None:
                return None
            buf = self._staging.get(name)
            if buf is None or buf.shape != host.shape:
                buf = self._staging[name] = torch.empty(host.shape, device=self.device, dtype=host.dtype)
            buf.copy_(host, non_blocking=True)
            return buf
        lat, ctx, ctxn = stage("lat", latents_host), stage("ctx", context_host), stage("ctxn", context_null_host)
        graph_mode, self.use_step_graph = self.use_step_graph, False      # the prompt is re-uploaded every call here: nothing step-invariant to replay
        try:
            out = self.step(lat, i, ctx, ctxn, y=y, freqs=freqs)
        finally:
            self.use_step_graph = graph_mode
        if out is None:
            return None
        latents_host.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return latents_host

    @torch.no_grad()
    def generate(self, context, context_null, latent_shape, seed=0, y=None, callback=None, decode=True, device_frames=False):
        """Full schedule: noise -> denoise loop -> VAE decode; returns {"x": uint8 CPU [3,F,H,W]} or latents, None if aborted
        (contract of WanAny2V.generate, any2video.py:1810-1826).  device_frames: return the clamped fp32 frames on the device instead
        (dist.generate_batch fuses their quantisation with the all-gather)."""
        g = torch.Generator(device="cpu").manual_seed(seed)
        latents = torch.randn(1, *latent_shape, dtype=f32, generator=g).to(self.device)      # any2video.py:1470
        context, context_null = context.to(self.device), (None if context_null is None else context_null.to(self.device))
        for i in range(self.num_steps):
            if self.step(latents, i, context, context_null, y=y, callback=callback) is None:
                return None
        if not decode or self.vae is None:
            return {"latents": latents}
        if device_frames:
            return {"x": self.vae.decode([latents[0]], 0)[0]}
        return {"x": self.vae.decode_to_cpu_uint8([latents[0]], 0)[0]}


class HunyuanDenoiser:
    """Denoise-loop call site of the Hunyuan Video 1.5 pipeline (models/hyvideo/diffusion/pipelines/pipeline_hunyuan_video.py
    :1597-1763): latent_model_input = cat(latents, cond_latents) (:1640-1650), transformer(cond) / transformer(uncond)
    (:1655/:1687), CFG combine (:1719-1743) and the flow-matching scheduler step (:1755; FlowMatchDiscreteScheduler, reverse=True, euler), here the
    Euler update fused with the combine.  guidance 6.0 / shift 9 are defaults/hunyuan_1_5_t2v.json."""

    def __init__(self, model, num_steps=30, shift=9.0, guide_scale=6.0, device="cuda", cfg_group=None, cfg_rank=0):
        self.model, self.device, self.guide_scale = model, torch.device(device), guide_scale
        # CFG-pair split (BASELINE configs[3]: Hunyuan 1.5 on 4 GPUs = 2 samples x 2 branches): rank 0 of the pair
        # runs the conditional forward, rank 1 the unconditional one, one 2-rank all-gather of the fp32 prediction per step, then both
        # apply the same fused combine + Euler update (deterministic: the latents stay replicated) -- as WanDenoiser does.
        self.cfg_group, self.cfg_rank = cfg_group, cfg_rank
        self.timesteps = flow_match_timesteps(num_steps, shift)              # scheduler.step = x + v (sigma_next - sigma) (:237-240)
        self.num_steps = num_steps
        self._interrupt = False

    @torch.no_grad()
    def step(self, latents, cond_latents, i, text, text_mask, text_null, text_null_mask, byt5=None, byt5_mask=None, freqs=None,
             text_states_2=None, guidance=None, byt5_null=None, byt5_null_mask=None, text_states_2_null=None, cfg_star=False,
             joint_pass=False, callback=None, pipeline=None, vision_states=None):
        """latents fp32 [1,C,T,H,W] (updated in place); cond_latents fp32 [1,C2,T,H,W] (Hunyuan 1.5 concat mask/cond channels) or
        None; text_null=None => no CFG (guidance-distilled HunyuanVideo 1.0: one forward with the guidance embedding).
        `byt5_null*` / `text_states_2_null`: the negative branch's glyph / pooled states (default: the positive ones);
        cfg_star = the CFG-Zero* rescale of the unconditional prediction (pipeline_hunyuan_video.py:1721-1731); joint_pass = both
        branches in ONE forward of batch 2 (:1687-1715) instead of two forwards (:1655-1685) -- same arithmetic per sample;
        `pipeline` = the object whose `_interrupt` the model polls once per block (default: this denoiser); `vision_states` = the
        image-encoder tokens of Hunyuan 1.5 i2v, the same for both branches (:1681)."""
        t = self.timesteps[i]
        dt = (t - self.timesteps[i + 1]) / 1000.0
        x = latents if cond_latents is None else torch.cat([latents, cond_latents], 1)
        tt = torch.tensor([t], dtype=f32)
        fr = dict(freqs_cos=None if freqs is None else freqs[0], freqs_sin=None if freqs is None else freqs[1],
                  pipeline=self if pipeline is None else pipeline, step_no=i, guidance=guidance, callback=callback)
        if vision_states is not None:
            fr["vision_states"] = vision_states
        pos = dict(text_states=text, text_mask=text_mask, byt5_text_states=byt5, byt5_text_mask=byt5_mask, text_states_2=text_states_2)
        uncond = None
        if text_null is None:
            cond = self.model(x, tt, **pos, **fr)
            if cond is None:
                return None
        else:
            neg = dict(text_states=text_null, text_mask=text_null_mask, byt5_text_states=byt5 if byt5_null is None else byt5_null,
                       byt5_text_mask=byt5_mask if byt5_null is None else byt5_null_mask,
                       text_states_2=text_states_2 if text_states_2_null is None else text_states_2_null)
            if self.cfg_group is not None:
                mine = self.model(x, tt, x_id=1 - self.cfg_rank, **(pos if self.cfg_rank == 0 else neg), **fr)
                pair = exchange_cfg_pair(mine, latents, self.cfg_group)
                if pair is None:
                    self._interrupt = True
                    return None
                cond, uncond = pair
            elif joint_pass:                                 # [uncond, cond] stacked along the batch, as the reference stacks them (:1445-1452)
                both = {k: (None if pos[k] is None else torch.cat([neg[k].to(pos[k].device), pos[k]], 0)) for k in pos}
                g2 = None if guidance is None else guidance.reshape(-1)[:1].repeat(2)
                ret = self.model(torch.cat([x, x], 0), tt.repeat(2), **both, **dict(fr, guidance=g2))
                if ret is None:
                    return None
                uncond, cond = ret[0:1].contiguous(), ret[1:2].contiguous()
            else:
                uncond = self.model(x, tt, x_id=0, **neg, **fr)          # the reference runs the unconditional branch first (j = 0)
                if uncond is None:
                    return None
                cond = self.model(x, tt, x_id=1, **pos, **fr)
                if cond is None:
                    return None
                uncond = uncond.contiguous()
        ops.cfg_euler_step_(latents, cond.contiguous(), uncond, self.guide_scale, dt, cfg_star=bool(cfg_star) and uncond is not None)
        return latents
