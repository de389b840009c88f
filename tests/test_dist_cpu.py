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
