"""CPU (gloo, world_size 2): host-side logic of the batch-split path -- sharding and the final frame all-gather."""
import os

import pytest
import torch
import torch.multiprocessing as mp

from wan2gp_b200 import dist as wd
from wan2gp_b200.pipeline import euler_timesteps


def test_shard():
    for n in range(0, 20):
        for w in (1, 2, 3, 4, 8):
            parts = [wd.shard(n, r, w) for r in range(w)]
            assert sum(parts, []) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def _worker(rank, world, port, n_samples, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    r, w = wd.init(backend="gloo")
    assert (r, w) == (rank, world)
    mine = wd.shard(n_samples, rank, world)
    # "decoded frames" of sample i are a deterministic function of i
    frames = torch.stack([torch.full((3, 5, 8, 8), i, dtype=torch.uint8) for i in mine], 0) if mine else torch.empty((0, 3, 5, 8, 8), dtype=torch.uint8)
    allf = wd.allgather_frames(frames)
    q.put((rank, allf.shape[0], [int(allf[i, 0, 0, 0, 0]) for i in range(allf.shape[0])]))
    torch.distributed.destroy_process_group()


@pytest.mark.parametrize("n_samples", [2, 3, 5])
def test_allgather_frames_gloo(n_samples):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + n_samples
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_samples, q)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in procs]
    [p.join(timeout=60) for p in procs]
    for rank, n, ids in res:
        assert n == n_samples and ids == list(range(n_samples))


def test_euler_schedule_matches_reference_formula():
    """shared/utils/euler_scheduler.py:35-51 with shift 12, 50 steps (defaults/t2v_2_2.json flow_shift)."""
    ts = euler_timesteps(50, 12.0)
    assert len(ts) == 51 and abs(ts[0] - 1000.0) < 1e-3 and ts[-1] == 0.0
    assert all(a > b for a, b in zip(ts, ts[1:]))
    t = 1000.0 - 999.0 / 49.0          # second linspace point
    assert abs(ts[1] - 12 * (t / 1000) / (1 + 11 * (t / 1000)) * 1000) < 1e-2
    assert sum(1 for v in ts[:-1] if v > 875) >= 1       # both Wan2.2 experts are used


class _FakeModel:
    """Stands in for WanModel on CPU: a deterministic function of (latents, context)."""

    def __call__(self, x, t, context, **kw):
        return [xi * 0.5 + ci.mean() * 0.1 + float(t[0]) * 1e-4 for xi, ci in zip(list(x), context)]


def _cfg_worker(rank, world, port, q):
    import torch.distributed as dist
    from oracle import wan_oracle
    from wan2gp_b200.pipeline import WanDenoiser
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    wd.init(backend="gloo")
    group, cfg_rank, sample, n_pairs = wd.make_cfg_pairs()
    g = torch.Generator().manual_seed(100 + sample)
    lat = torch.randn(1, 16, 2, 4, 4, generator=g)
    ctx, ctxn = torch.randn(1, 8, 16, generator=g), torch.zeros(1, 8, 16)
    den = WanDenoiser(_FakeModel(), num_steps=4, shift=5.0, guide_scale=4.0, device="cpu", cfg_group=group, cfg_rank=cfg_rank)
    den._combine_step = lambda latents, c, u, gs, dt, star: latents.sub_(dt * wan_oracle.cfg_combine(c, u, gs))
    ref = WanDenoiser(_FakeModel(), num_steps=4, shift=5.0, guide_scale=4.0, device="cpu")
    ref._combine_step = den._combine_step
    a, b = lat.clone(), lat.clone()
    for i in range(4):
        den.step(a, i, ctx, ctxn)
        ref.step(b, i, ctx, ctxn)
    q.put((rank, sample, n_pairs, bool(torch.allclose(a, b, atol=1e-6)), float(a.sum())))
    dist.destroy_process_group()


def test_cfg_pair_split_gloo():
    """4 ranks = 2 CFG pairs: each rank runs one branch, the pair all-gathers the prediction, latents stay replicated and equal
    the single-process joint-pass result."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_cfg_worker, args=(r, 4, 29641, q)) for r in range(4)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=180) for _ in procs)
    [p.join(timeout=60) for p in procs]
    assert [r[1] for r in res] == [0, 0, 1, 1] and all(r[2] == 2 and r[3] for r in res)
    assert res[0][4] == res[1][4] and res[2][4] == res[3][4] and res[0][4] != res[2][4]


class _AbortingModel(_FakeModel):
    """Returns [None] (the reference's abort contract, model.py:1995-1998) from step `at` on -- on ONE rank only."""

    def __init__(self, abort):
        self.abort, self.calls = abort, 0

    def __call__(self, x, t, context, **kw):
        self.calls += 1
        if self.abort and self.calls >= 2:
            return [None] * len(list(x))
        return super().__call__(x, t, context, **kw)


def _abort_worker(rank, world, port, q):
    import torch.distributed as dist
    from wan2gp_b200.pipeline import WanDenoiser
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    wd.init(backend="gloo")
    group, cfg_rank, sample, _ = wd.make_cfg_pairs()
    lat = torch.randn(1, 16, 2, 4, 4, generator=torch.Generator().manual_seed(7))
    ctx, ctxn = torch.randn(1, 8, 16), torch.zeros(1, 8, 16)
    den = WanDenoiser(_AbortingModel(abort=(rank == 1)), num_steps=4, shift=5.0, device="cpu", cfg_group=group, cfg_rank=cfg_rank)
    den._combine_step = lambda latents, c, u, gs, dt, star: latents.sub_(dt * (u + gs * (c - u)))
    outcomes = [den.step(lat, i, ctx, ctxn) is None for i in range(3)]
    # second part: generate_batch where only rank 1's denoiser aborts
    class _Den:
        def generate(self, context, context_null, latent_shape, seed=0, device_frames=False):
            return None if rank == 1 else {"x": torch.zeros(3, 5, 16, 16, dtype=torch.uint8)}
    res = wd.generate_batch(lambda: _Den(), [None, None], None, (16, 2, 2, 2), [0, 1], device="cpu", fused=False)
    q.put((rank, outcomes, res is None))
    dist.destroy_process_group()


def test_interrupt_is_agreed_collectively_gloo():
    """`_interrupt` is per process.  The partner of an interrupted rank must not be left waiting in the pair's all-gather
    (pipeline.WanDenoiser.step) nor in the final frame all-gather (dist.generate_batch): both ranks return None together."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_abort_worker, args=(r, 2, 29655, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in procs)
    [p.join(timeout=60) for p in procs]
    for rank, outcomes, batch_none in res:
        assert outcomes == [False, True, True], (rank, outcomes)      # step 0 completes on both, step 1 aborts on BOTH
        assert batch_none


class _FakeHYModel:
    """Stands in for HYVideoDiffusionTransformer on CPU: a deterministic function of (x, text states); `fail_from` makes the forward
    return None (the reference's abort contract, models.py:1146-1149) from that call on."""
    out_channels = 8

    def __init__(self, fail_from=None):
        self.calls, self.fail_from = 0, fail_from

    def __call__(self, x, t, text_states=None, text_mask=None, **kw):
        self.calls += 1
        if self.fail_from is not None and self.calls >= self.fail_from:
            return None
        return x[:, :8] * 0.5 + text_states.float().mean() * 0.1 + float(t[0]) * 1e-4


def _hy_cfg_worker(rank, world, port, q):
    import torch.distributed as dist
    import wan2gp_b200.pipeline as pl
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    wd.init(backend="gloo")
    group, cfg_rank, sample, n_pairs = wd.make_cfg_pairs()
    pl.ops.cfg_euler_step_ = lambda lat, c, u, g, dt, cfg_star=False: lat.sub_(dt * (c if u is None else u + g * (c - u)))   # CPU stand-in of the fused kernel
    g = torch.Generator().manual_seed(200 + sample)
    lat = torch.randn(1, 8, 2, 4, 4, generator=g)
    cond_lat = torch.zeros(1, 9, 2, 4, 4)
    txt, txtn = torch.randn(1, 6, 16, generator=g), torch.randn(1, 6, 16, generator=g)
    tm = torch.ones(1, 6, dtype=torch.long)
    den = pl.HunyuanDenoiser(_FakeHYModel(), num_steps=4, shift=9.0, guide_scale=6.0, device="cpu", cfg_group=group, cfg_rank=cfg_rank)
    ref = pl.HunyuanDenoiser(_FakeHYModel(), num_steps=4, shift=9.0, guide_scale=6.0, device="cpu")
    a, b = lat.clone(), lat.clone()
    for i in range(4):
        assert den.step(a, cond_lat, i, txt, tm, txtn, tm) is a
        assert ref.step(b, cond_lat, i, txt, tm, txtn, tm) is b
    same, calls = bool(torch.allclose(a, b, atol=1e-6)), den.model.calls
    # abort on ONE rank of pair 0 only: its partner leaves the step with it, the other pair finishes
    den2 = pl.HunyuanDenoiser(_FakeHYModel(fail_from=2 if rank == 1 else None), num_steps=4, shift=9.0, guide_scale=6.0, device="cpu",
                              cfg_group=group, cfg_rank=cfg_rank)
    c = lat.clone()
    outcomes = [den2.step(c, cond_lat, i, txt, tm, txtn, tm) is None for i in range(2)]
    q.put((rank, sample, n_pairs, same, calls, float(a.sum()), outcomes))
    dist.destroy_process_group()


def test_hunyuan_cfg_pair_split_gloo():
    """BASELINE configs[3] (Hunyuan 1.5 on 4 GPUs) as 2 samples x 2 CFG branches: each rank runs ONE forward per step, the pair exchanges the
    prediction, latents stay replicated and equal the single-process two-forward result; an abort on one rank is agreed inside its pair."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_hy_cfg_worker, args=(r, 4, 29671, q)) for r in range(4)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=180) for _ in procs)
    [p.join(timeout=60) for p in procs]
    assert [r[1] for r in res] == [0, 0, 1, 1] and all(r[2] == 2 and r[3] for r in res)
    assert all(r[4] == 4 for r in res)                                   # one forward per step and rank (the unsplit denoiser runs two)
    assert res[0][5] == res[1][5] and res[2][5] == res[3][5] and res[0][5] != res[2][5]
    assert res[0][6] == [False, True] and res[1][6] == [False, True]     # pair 0 (ranks 0, 1): step 0 completes, step 1 aborts on BOTH
    assert res[2][6] == [False, False] and res[3][6] == [False, False]   # pair 1 is unaffected
